"""Drop-in for the reference's ``Robustness.py``: post-integration reductions on the
heatmaps written by ``BCB_latest.parallel_routine``.

``robustness`` / ``routine`` keep the reference's arithmetic (Robustness.py:25-52) on the
loaded arrays; ``robustness_map`` and ``ensemble`` are the fused device versions: the whole
R[beta][T] map with the per-column first-maximum search in one pass, and the sampled
generalisation of BASELINE config 4 (noise on gamma and T, thesis p.29).
"""

import sys

import numpy as np

from . import engine

DATA_TYPE = ["_probability_static.npy", "_probability_lin.npy", "_probability_sqrt.npy",
             "_probability_cbrt.npy", "_probability_cerf_3.npy"]


def load_data_file(dim, flag):
    """Robustness.py:11-23 -- same file names."""
    probability = np.load(str(dim) + DATA_TYPE[flag])
    time = np.load(str(dim) + '_circ_time.npy')
    beta = np.load(str(dim) + '_circ_beta.npy')
    return probability, time, beta


def robustness(probability, beta_index, time_index, variation):
    """R = ((p0 - p[beta+v]) + (p0 - p[beta-v])) / 2, 4 decimals (Robustness.py:25-31)."""
    R_pos = probability[beta_index, time_index] - probability[beta_index + variation, time_index]
    R_neg = probability[beta_index, time_index] - probability[beta_index - variation, time_index]
    return round(float((R_pos + R_neg) / 2), 4)


def robustness_map(probability, variation, device=None):
    """Unrounded R for every cell and the first-maximum beta index of every time column,
    computed on the device."""
    return engine.robustness_map(probability, variation, device=device)


def routine(dim, flag):
    """Robustness.py:34-52: argmax over beta at time column 2, R at 2 and 4 cells."""
    probability, time, beta = load_data_file(dim, flag)
    R2, arg = robustness_map(probability, 2)
    max_index = int(arg[2])
    if max_index < 0:        # no p > 0 in that column: the reference's scan leaves max_index unset
        raise ValueError("no positive probability in time column 2 of the %s heatmap" % (dim,))
    out = (dim, robustness(probability, max_index, 2, 2), robustness(probability, max_index, 2, 4))
    print(*out)
    return out


def ensemble(problem, T, gamma, thresholds=(0.8, 0.9, 0.95, 0.99), p_reference=None,
             device=None):
    """Noisy-(T, gamma) ensemble: integrate every realisation and reduce on the device.
    Returns mean / std / min / max / fraction >= threshold and, if the unperturbed
    ``p_reference`` is given, the sampled robustness R = p_reference - mean."""
    import torch
    dev = engine._device(device)
    dT = torch.as_tensor(np.asarray(T, dtype=np.float64)).to(dev)
    dG = torch.as_tensor(np.asarray(gamma, dtype=np.float64)).to(dev)
    p, norm = engine.rk4_batch(problem, dT, dG, device=dev)
    stats = engine.ensemble_stats(p, thresholds, device=dev)
    if p_reference is not None:
        stats["R_sampled"] = float(p_reference) - stats["mean"]
    return stats


if __name__ == '__main__':
    flag = int(sys.argv[1])
    for dim in np.arange(7, 71 + 1, 2):
        routine(int(dim), flag)
