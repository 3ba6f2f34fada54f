"""Parity at BASELINE.json's full sizes, through the GRID api (so the machine-filling kernel
geometry is the one tested), against the C oracle on the same inputs.

  C2  cycle N=64, the real 100 x 100 (T, gamma) grid, lin / cerf_3 / static: every cell, flat
      argmax, per-column argmax, tau-argmin (BCB.py:239-259, Robustness.py:41-45 first-max scan)
  C3  complete N=4096, gamma on the Laplacian: strided 64 x 64 sub-grid of the real 256 x 256
      grid with argmax, one full T row, one full gamma column
  C4  cycle N=256: all 10^5 noisy realisations, p per point and the device statistics
  C5  cycle N=1024, S=4096: strided 64 x 64 sub-grid of the real 1024 x 1024 grid with argmax,
      one full T row, one full gamma column; N=65536 streaming variant: the 32-point diagonal
      of its 32 x 32 grid

Tolerances (north star): |dp| <= 1e-10 per success probability, argmax indices identical.
The oracle side is ~1 minute of host time on 16 threads.
"""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import rk4_numpy as rk

pytestmark = pytest.mark.gpu

P_TOL = 1e-10


def _tau_argmin(p, t, t_min):
    tau = np.where(t[None, :] >= t_min, t[None, :] / p, np.inf)
    return np.unravel_index(np.argmin(tau), tau.shape), tau.min()


def _check_heatmap(qw, p, norm, po, no, t, n):
    assert p.shape == po.shape
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    assert np.argmax(p) == np.argmax(po)                            # flat first maximum
    assert (np.argmax(p, axis=0) == np.argmax(po, axis=0)).all()    # per T column (ROB:41-45)
    t_min = (np.pi / 2) * np.sqrt(n)
    (jo, io), tau_o = _tau_argmin(po, t, t_min)
    s = qw.heatmap_summary(p, t, t_min=t_min)                       # the device reductions
    assert s["argmax"] == tuple(int(x) for x in np.unravel_index(np.argmax(po), po.shape))
    assert s["tau_argmin"] == (int(jo), int(io))
    assert s["tau_min"] == pytest.approx(tau_o, rel=1e-9)
    _, arg = qw.robustness_map(p, 2)
    for i in range(p.shape[1]):
        assert arg[i] == rk.first_argmax(po[:, i])


@pytest.mark.parametrize("sched", ["lin", "cerf_3", "static"])
def test_c2_full_grid(qw, sched):
    n, S = 64, 1024
    t, b = np.linspace(0.1, 64, 100), np.linspace(0, 2.5, 100)
    prob = qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S)
    assert qw.plan(prob)["path"].startswith("packed")
    p, norm = qw.rk4_heatmap(prob, t, b)
    po, no = c_oracle.heatmap(n, "circular", sched, t, b, n_steps=S)
    _check_heatmap(qw, p, norm, po, no, t, n)
    # the symmetry-reduced default of the drop-in modules against the same full-ring oracle
    pm, nm = qw.rk4_heatmap(qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S,
                                       flags=8192), t, b)
    _check_heatmap(qw, pm, nm, po, no, t, n)


def test_c2_trio_one_launch(qw):
    """lin, cerf_3 and static as ONE multi-problem launch: bitwise the three single launches."""
    if not hasattr(qw, "rk4_heatmaps"):
        pytest.skip("multi-problem launch not built")
    n, S = 64, 1024
    t, b = np.linspace(0.1, 64, 100), np.linspace(0, 2.5, 100)
    probs = [qw.Problem(n=n, topology="circular", schedule=s, n_steps=S)
             for s in ("lin", "cerf_3", "static")]
    multi = qw.rk4_heatmaps(probs, [t] * 3, [b] * 3)
    for prob, (pm, nm) in zip(probs, multi):
        p, norm = qw.rk4_heatmap(prob, t, b)
        assert np.array_equal(p, pm) and np.array_equal(norm, nm)


def test_c4_all_realisations(qw):
    n, S = 256, 512
    T0 = (np.pi / 2) * np.sqrt(n)
    gs = np.linspace(0, 2.5, 101)
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S)
    pc, _ = qw.rk4_batch(prob, np.full(101, T0), gs)
    pco, _, _ = c_oracle.rk4_batch(n, "circular", "lin", np.full(101, T0), gs, n_steps=S)
    assert np.argmax(pc) == np.argmax(pco) and np.abs(pc - pco).max() <= P_TOL
    g0 = gs[int(np.argmax(pco))]
    rng = np.random.default_rng(20200912)
    g = np.clip(g0 * (1 + 0.05 * rng.standard_normal(100000)), 0, None)
    T = np.clip(T0 * (1 + 0.05 * rng.standard_normal(100000)), 0, None)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(n, "circular", "lin", T, g, n_steps=S)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    pm, nm = qw.rk4_batch(qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S,
                                     flags=8192), T, g)            # half-ring default of the drop-ins
    assert np.abs(pm - po).max() <= P_TOL and np.abs(nm - no).max() <= P_TOL
    thr = (0.8, 0.9, 0.95, 0.99)
    thr2 = tuple(float(x) for x in np.quantile(po, [0.1, 0.5, 0.9, 0.99]))
    for th in (thr, thr2):
        st = qw.ensemble_stats(p, thresholds=th)
        assert st["mean"] == pytest.approx(po.mean(), abs=1e-11)
        assert st["std"] == pytest.approx(po.std(), abs=1e-11)
        assert st["min"] == pytest.approx(po.min(), abs=P_TOL)
        assert st["max"] == pytest.approx(po.max(), abs=P_TOL)
        for x, f in st["fraction_ge"].items():
            # a realisation within 1e-10 of a threshold may fall on either side
            near = int((np.abs(po - x) <= P_TOL).sum())
            assert abs(f * po.size - (po >= x).sum()) <= near + 1e-6
    assert pc[int(np.argmax(pco))] - st["mean"] == pytest.approx(pco.max() - po.mean(), abs=1e-11)


def test_c4_localization_sample(qw):
    """The second ensemble of config 4 (T0 = N^2/2, S = 131072) on 32 realisations."""
    n, S = 256, 131072
    rng = np.random.default_rng(20200913)
    g = np.clip(1.0 * (1 + 0.05 * rng.standard_normal(32)), 0, None)
    T = np.clip(0.5 * n * n * (1 + 0.05 * rng.standard_normal(32)), 0, None)
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(n, "circular", "lin", T, g, n_steps=S)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL


def _strided_grid_checks(qw, kw, okw, t, b, stride, row, col):
    n = kw["n"]
    prob = qw.Problem(**kw)
    ts, bs = t[::stride], b[::stride]
    p, norm = qw.rk4_heatmap(prob, ts, bs)
    po, no = c_oracle.heatmap(n, okw["topology"], okw["schedule"], ts, bs,
                              n_steps=kw["n_steps"], gamma_on=kw.get("gamma_on", "oracle"))
    _check_heatmap(qw, p, norm, po, no, ts, n)
    # one full T row and one full gamma column of the real grid, as heatmap calls
    pr, nr = qw.rk4_heatmap(prob, t, b, rows=(row, row + 1))
    pro, nro = c_oracle.heatmap(n, okw["topology"], okw["schedule"], t, b[row:row + 1],
                                n_steps=kw["n_steps"], gamma_on=kw.get("gamma_on", "oracle"))
    assert np.abs(pr - pro).max() <= P_TOL and np.abs(nr - nro).max() <= P_TOL
    assert np.argmax(pr) == np.argmax(pro)
    pcol, ncol = qw.rk4_heatmap(prob, t[col:col + 1], b)
    pcolo, ncolo = c_oracle.heatmap(n, okw["topology"], okw["schedule"], t[col:col + 1], b,
                                    n_steps=kw["n_steps"], gamma_on=kw.get("gamma_on", "oracle"))
    assert np.abs(pcol - pcolo).max() <= P_TOL and np.abs(ncol - ncolo).max() <= P_TOL
    assert np.argmax(pcol) == np.argmax(pcolo)
    return po


def test_c5_strided_subgrid_row_and_column(qw):
    n = 1024
    t, b = np.linspace(0.1, 1024, 1024), np.linspace(0, 2.5, 1024)
    kw = dict(n=n, topology="circular", schedule="lin", n_steps=4096)
    assert qw.plan(qw.Problem(**kw))["path"] == "horner-cycle"
    po = _strided_grid_checks(qw, kw, kw, t, b, 16, row=410, col=300)
    # the half-ring reduction (default of the drop-in modules) against the FULL-ring oracle
    pm, nm = qw.rk4_heatmap(qw.Problem(flags=8192, **kw), t[::16], b[::16])
    assert np.abs(pm - po).max() <= P_TOL and np.argmax(pm) == np.argmax(po)


def test_c3_strided_subgrid_row_and_column(qw):
    n = 4096
    t, b = np.linspace(0.1, 512, 256), np.linspace(0.5 / n, 2.5 / n, 256)
    kw = dict(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian", n_steps=2048)
    assert qw.plan(qw.Problem(**kw))["path"] == "horner-complete"
    po = _strided_grid_checks(qw, kw, kw, t, b, 4, row=100, col=200)
    pm, nm = qw.rk4_heatmap(qw.Problem(flags=8192, **kw), t[::4], b[::4])     # pair reduction
    assert np.abs(pm - po).max() <= P_TOL and np.argmax(pm) == np.argmax(po)


def test_c5_streaming_variant_diagonal(qw):
    """cycle N=65536, S=256: the 32 diagonal points of the 32 x 32 grid of the streaming
    variant, default steps-per-pass and the plain one-step-per-pass stream."""
    n, S = 65536, 256
    T, g = np.linspace(0.1, 64, 32), np.linspace(0, 2.5, 32)
    po, no, _ = c_oracle.rk4_batch(n, "circular", "lin", T, g, n_steps=S)
    for flags in (0, 8, 8192):
        prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S, flags=flags)
        p, norm = qw.rk4_batch(prob, T, g)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
        assert np.argmax(p) == np.argmax(po)
