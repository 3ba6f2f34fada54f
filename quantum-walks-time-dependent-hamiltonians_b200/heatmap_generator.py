"""Drop-in for the reference's ``heatmap_generator.py``: the difference of two saved
heatmaps, rendered with a log colour norm (heatmap_generator.py:22-71).  Rendering is
outside the accelerated path and needs matplotlib, which this image does not ship; the
numeric part (``difference``) has no such dependency."""

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


def render(arr, time_array, beta_array, *extra, dimension=None, file_name=None, **kw):
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    time_array = [round(float(t), 1) for t in time_array]
    beta_array = [round(float(b), 2) for b in beta_array]
    plt.imshow(arr, aspect=1., origin='lower', cmap='inferno_r')
    plt.tick_params(axis='both', which='major', labelsize=7)
    plt.xticks(range(len(time_array)), time_array, rotation='vertical')
    plt.yticks(range(len(beta_array)), beta_array)
    plt.xlabel('Time', fontweight="bold")
    plt.ylabel('Beta', fontweight="bold")
    plt.title('Probability N = ' + str(dimension), y=1.04, fontweight="bold", ha='center')
    plt.colorbar()
    plt.savefig(file_name or (str(dimension) + '_heatmap.pdf'))
    plt.clf()
    plt.close()


def heatmap2d(arr, time_array, beta_array):
    """heatmap_generator.py:22-55."""
    try:
        import matplotlib  # noqa: F401
    except Exception:
        print("heatmap2d: matplotlib is not installed; plot skipped")
        return None
    return render(arr, time_array, beta_array, dimension=dimension,
                  file_name=str(dimension) + '_heatmap_sqrt.pdf')


if __name__ == '__main__':
    dimension = int(sys.argv[1])
    p, time_array, beta_array = difference(dimension)
    heatmap2d(p, time_array, beta_array)
