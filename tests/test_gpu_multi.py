"""Many small problems in one grid (qw_rk4_multi_grid_dev) and the warp-packed half ring
(QW_FLAG_MIRROR below 160 sites): the shapes of the reference's production sweep
(BCB_latest.py:309-320: odd dimensions 3 .. 71, 75 beta x 4 T, four schedules)."""
import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu

P_TOL, PSI_TOL = 1e-10, 1e-11
MIRROR = 8192


@pytest.mark.parametrize("n", [6, 7, 8, 9, 15, 16, 17, 33, 51, 64, 71, 100, 128, 159])
def test_packed_half_ring_vs_full_ring_oracle(qw, n):
    rng = np.random.default_rng(n)
    P = 301
    T = rng.uniform(0.0, 10.0, P)
    g = rng.uniform(0.0, 2.5, P)
    T[::37] = 0.0
    for sched, gamma_on, target, S in (("cbrt", "oracle", None, 70), ("cerf_5", "laplacian", 1, 33)):
        prob = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on,
                          target=target, n_steps=S, flags=MIRROR)
        assert qw.plan(prob)["path"] == "packed-mirror-cycle"
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, n_steps=S,
                                             gamma_on=gamma_on, target=target, return_psi=True)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched)
    # fixed step size on a heatmap: a warp takes rows of one T column
    t, b = np.array([0.0, 0.31, 1.7, 0.9]), np.linspace(0.1, 2.4, 11)
    prob = qw.Problem(n=n, topology="circular", schedule="sqrt", step_size=1e-2, flags=MIRROR)
    p, norm = qw.rk4_heatmap(prob, t, b)
    po, no = c_oracle.heatmap(n, "circular", "sqrt", t, b, step_size=1e-2)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL


@pytest.mark.parametrize("flags", [0, MIRROR])
def test_multi_problem_launch_is_bitwise_the_single_launches(qw, flags):
    beta = np.array([round(0.05 * k, 2) for k in range(2, 80) if k not in (21, 41, 61)])
    probs, ts, bs = [], [], []
    for sched in ("lin", "cerf_3", "static"):
        for n in (3, 5, 7, 9, 15, 16, 31, 51, 64, 71):
            probs.append(qw.Problem(n=n, topology="circular", schedule=sched, step_size=5e-3,
                                    flags=flags))
            ts.append(np.array([np.pi / 4, 1.0, np.pi / 2, 2.0]) * np.sqrt(n))
            bs.append(beta)
    # one problem of another family rides along and is launched on its own
    probs.append(qw.Problem(n=300, topology="complete", schedule="lin", n_steps=40))
    ts.append(np.linspace(0.001, 0.01, 3))
    bs.append(np.linspace(100.0, 400.0, 5))
    before = qw.launch_count()
    multi = qw.rk4_heatmaps(probs, ts, bs)
    launches = qw.launch_count() - before
    assert launches <= 10                             # 31 problems, a handful of grids
    for prob, t, b, (pm, nm) in zip(probs, ts, bs, multi):
        p, norm = qw.rk4_heatmap(prob, t, b)
        assert np.array_equal(p, pm) and np.array_equal(norm, nm), (prob.n, prob.schedule)
    # and against the oracle for one of them
    k = 7
    po, no = c_oracle.heatmap(probs[k].n, "circular", "lin", ts[k], bs[k], step_size=5e-3)
    assert np.abs(multi[k][0] - po).max() <= P_TOL


def test_multi_problem_launch_n_steps_rule_and_device_tensors(qw):
    import torch
    dev = torch.device("cuda")
    t = torch.linspace(0.1, 64, 50, dtype=torch.float64, device=dev)
    b = torch.linspace(0, 2.5, 40, dtype=torch.float64, device=dev)
    probs = [qw.Problem(n=64, topology="circular", schedule=s, n_steps=256)
             for s in ("lin", "cerf_3", "static", "sqrt")]
    before = qw.launch_count()
    out = qw.rk4_heatmaps(probs, [t] * 4, [b] * 4)
    assert qw.launch_count() - before == 1            # one grid for the four schedules
    for prob, (pm, nm) in zip(probs, out):
        assert pm.is_cuda
        p, norm = qw.rk4_heatmap(prob, t, b)
        assert torch.equal(p, pm) and torch.equal(norm, nm)
    assert qw.rk4_heatmaps([], [], []) == []


@pytest.mark.parametrize("topo,n", [("circular", 3), ("circular", 5), ("circular", 16),
                                    ("circular", 31), ("circular", 32), ("complete", 1),
                                    ("complete", 2), ("complete", 15), ("complete", 32)])
def test_parallel_in_time_vs_oracle(qw, topo, n):
    """A handful of points on a tiny graph: the time axis is cut into chunk propagators
    (qw_pit.cu).  Every schedule, both step rules, custom target, psi, T = 0."""
    rng = np.random.default_rng(n)
    scale = 1.0 if topo == "circular" else 1.0 / max(n, 1)
    T = np.array([0.0, 3.0, 16.0, 7.7, 0.4]) * scale
    g = np.array([1.0, 0.3, 1.0, 2.2, 1.7]) * (1.0 if topo == "circular" else n)
    for sched in ("lin", "sqrt", "cbrt", "cerf_3", "cerf_5", "cerf_7", "static"):
        for kw in (dict(n_steps=1024), dict(n_steps=130), dict(step_size=2e-3 * scale)):
            target = None if sched != "cbrt" else min(1, n - 1)
            prob = qw.Problem(n=n, topology=topo, schedule=sched, target=target, **kw)
            p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)       # host entry: any rule
            po, no, psio, _ = c_oracle.rk4_batch(n, topo, sched, T, g, target=target,
                                                 return_psi=True, **kw)
            assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL, (sched, kw)
            assert np.abs(psi - psio).max() <= PSI_TOL, (sched, kw)
    prob = qw.Problem(n=n, topology=topo, schedule="lin", n_steps=1024)
    assert qw.plan(prob)["path"] in ("parallel-in-time", "packed-cycle", "packed-any-cycle",
                                     "packed-complete", "generic", "resident-complete")
    # the warp-packed / resident kernels on the same points (QW_FLAG_NO_PACK keeps PIT off)
    pk, nk = qw.rk4_batch(qw.Problem(n=n, topology=topo, schedule="lin", n_steps=1024, flags=512),
                          T, g)
    pp, np_ = qw.rk4_batch(prob, T, g)
    assert np.abs(pk - pp).max() <= 1e-12 and np.abs(nk - np_).max() <= 1e-12


def test_config1_value_and_host_entry(qw):
    prob = qw.Problem(n=16, topology="circular", schedule="lin", n_steps=1024)
    p, norm = qw.rk4_batch(prob, [16.0], [1.0])
    assert abs(p[0] - 0.42038739054526) < 1e-10
    big_T = np.linspace(0.1, 16, 5000)
    pb, nb = qw.rk4_batch(prob, big_T, np.full(5000, 1.0))   # staged copies, throughput kernel
    po, no, _ = c_oracle.rk4_batch(16, "circular", "lin", big_T, np.full(5000, 1.0), n_steps=1024)
    assert np.abs(pb - po).max() <= P_TOL and np.abs(nb - no).max() <= P_TOL


def test_fixed_geometry_makes_results_independent_of_the_cut(qw):
    """QW_FLAG_FIXED_GEOMETRY (32768): a point's result depends on the problem only -- a row
    block, a single row and a three-point call are bitwise the rows of the whole heatmap."""
    t, b = np.linspace(0.1, 40, 9), np.linspace(0, 2.5, 14)
    for n, topo in ((16, "circular"), (64, "circular"), (71, "circular"), (256, "circular"),
                    (30, "complete")):
        for flags in (32768, 32768 | MIRROR):
            bb = b if topo == "circular" else b * n
            tt = t if topo == "circular" else t / n
            prob = qw.Problem(n=n, topology=topo, schedule="cerf_3", n_steps=200, flags=flags)
            p, norm = qw.rk4_heatmap(prob, tt, bb)
            for rows in ((0, 14), (3, 9), (13, 14)):
                ps, ns = qw.rk4_heatmap(prob, tt, bb, rows=rows)
                assert np.array_equal(ps, p[rows[0]:rows[1]]) and np.array_equal(ns, norm[rows[0]:rows[1]])
            pb, nb = qw.rk4_batch(prob, tt[[1, 4, 8]], bb[[2, 2, 11]])
            assert np.array_equal(pb, p[[2, 2, 11], [1, 4, 8]])
    lib = qw._lib.load()
    assert lib.qw_scratch_trim() == 0
