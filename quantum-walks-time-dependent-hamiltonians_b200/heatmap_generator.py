"""Drop-in for the reference's ``heatmap_generator.py``: the difference of two saved
heatmaps (heatmap_generator.py:59-70).  The matplotlib rendering of heatmap_generator.py:22-55
is out of scope: ``heatmap2d`` keeps the signature and does nothing."""

import sys

import numpy as np

dimension = None


def difference(dim):
    """probability(lin) - probability(sqrt) and its axes (heatmap_generator.py:61-70)."""
    p1 = np.load(str(dim) + '_probability.npy')
    p = np.load(str(dim) + '_probability_sqrt.npy')
    beta_array = np.load(str(dim) + '_beta_array.npy')
    time_array = np.load(str(dim) + '_time_array.npy')
    return p1 - p, time_array, beta_array


def heatmap2d(arr, time_array, beta_array):
    """Signature of heatmap_generator.py:22-55; no rendering."""
    print("heatmap2d: plotting is outside the accelerated path (SURVEY.md section 2); "
          "use difference() for the numbers")
    return None


if __name__ == '__main__':
    dimension = int(sys.argv[1])
    p, time_array, beta_array = difference(dimension)
    heatmap2d(p, time_array, beta_array)
