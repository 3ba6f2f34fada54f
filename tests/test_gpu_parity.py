"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden
fixtures produced by the reference's own code.

Tolerances (BASELINE.json north_star): |dp| <= 1e-10 absolute per success probability,
argmax (T, gamma) indices identical.  psi is compared at 1e-11.
"""

import numpy as np
import pytest

from oracle import c_oracle, rk4_numpy as rk
from oracle.defs import default_target

pytestmark = pytest.mark.gpu

P_TOL = 1e-10
PSI_TOL = 1e-11


def _problem(qw, n, topo, sched, rule, **kw):
    return qw.Problem(n=n, topology=topo, schedule=sched, **rule, **kw)


def test_golden_rk4_over_reference_rhs(qw, golden):
    """Every RK4 trajectory driven by the reference's own schrodinger_equation."""
    meta, arr = golden
    for c in meta["rk4_L0"]:
        prob = _problem(qw, c["n"], c["topology"], c["schedule"], c["rule"])
        p, norm, psi = qw.rk4_batch(prob, [c["T"]], [c["beta"]], return_psi=True)
        assert abs(p[0] - c["p"]) <= P_TOL, c
        assert abs(norm[0] - c["norm"]) <= P_TOL, c
        assert np.abs(psi[0] - arr["L0_psi_%d" % c["key"]]).max() <= PSI_TOL, c


def test_golden_rhs(qw, golden):
    meta, arr = golden
    for c in meta["rhs"]:
        prob = qw.Problem(n=c["n"], topology=c["topology"], schedule=c["schedule"], n_steps=1)
        dy = qw.rhs(prob, c["t"], c["T"], c["beta"], arr["rhs_y_%d" % c["key"]])
        ref = arr["rhs_dy_%d" % c["key"]]
        assert np.abs(dy - ref).max() <= 1e-12 * np.abs(ref).max(), c


def test_golden_static_and_rk45_are_close(qw, golden):
    meta, _ = golden
    for c in meta["static"]:
        prob = qw.Problem(n=c["n"], topology="circular", schedule="static",
                          n_steps=c["n_steps"])
        p, _ = qw.rk4_batch(prob, [c["time"]], [c["gamma"]])
        assert abs(p[0] - c["p_rk4"]) <= P_TOL
        assert abs(p[0] - c["p_expm"]) <= 1e-8           # RK4 truncation vs exact expm
    for c in meta["rk45"]:
        prob = qw.Problem(n=c["n"], topology=c["topology"], schedule=c["schedule"],
                          n_steps=8192)
        p, _ = qw.rk4_batch(prob, [c["T"]], [c["beta"]])
        tol = 2e-4 if c["schedule"] in ("sqrt", "cbrt") else 3e-5
        assert abs(p[0] - c["p"]) <= tol                  # the reference as shipped (RK45)


@pytest.mark.parametrize("topo", ["circular", "complete"])
@pytest.mark.parametrize("n", [3, 5, 16, 17, 31, 32, 33, 51, 64, 71, 100, 255, 256, 1000, 1024])
def test_random_batches_vs_oracle(qw, topo, n):
    rng = np.random.default_rng(1000 * n + (topo == "complete"))
    scheds = ["lin", "sqrt", "cbrt", "cerf_3", "cerf_5", "cerf_7", "static"]
    for sched in scheds:
        P = 5
        if topo == "circular":
            T = rng.uniform(0.0, 12.0, P)
            g = rng.uniform(0.0, 2.5, P)
            S = 96
        else:                         # spectral radius ~ N: keep h * N small
            T = rng.uniform(0.0, 40.0 / n, P)
            g = rng.uniform(0.0, 2.0 * n, P)
            S = 160
        T[0] = 0.0                    # T == 0 guard
        prob = qw.Problem(n=n, topology=topo, schedule=sched, n_steps=S)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, topo, sched, T, g, n_steps=S, return_psi=True)
        assert np.abs(p - po).max() <= P_TOL, (n, topo, sched)
        assert np.abs(norm - no).max() <= P_TOL, (n, topo, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, topo, sched)
        assert p[0] == pytest.approx(1.0 / n, abs=1e-15)


@pytest.mark.parametrize("n,flags", [(48, 1), (1024, 1), (2050, 0), (1031, 0), (96, 4), (4096, 4),
                                     (1024, 256), (512, 256), (4096, 256), (1024, 256 | 32),
                                     (1024, 256 | 64), (1024, 256 | 128), (2048, 0), (3072, 0),
                                     (4096, 256 | 32), (512, 0), (528, 0), (16, 512), (64, 512),
                                     (256, 512), (96, 512), (8, 0), (24, 0), (40, 0), (56, 0),
                                     (112, 0), (192, 0), (224, 0), (4096, 1024), (4096, 256 | 1024),
                                     (96, 1024), (2048, 1024)])
def test_generic_and_exact_sum_paths(qw, n, flags):
    """flags: 1 = force the global-memory kernel, 4 = block-reduce sum(u) in every stage,
    256 = at most 8 sites per thread, 32 / 64 / 128 = first-generation / mbarrier / 4-blocks
    variants of the resident kernels, 512 = no warp packing for small rings; n = 2050 / 1031
    do not fit the resident kernels; 8 ... 224 exercise every R of the packed kernel; 1024 =
    barrier-free variant of the complete-graph kernel."""
    rng = np.random.default_rng(n)
    for topo in ("circular", "complete"):
        if (flags & 4) and topo == "circular":
            continue
        T = rng.uniform(1.0, 6.0, 3) / (1.0 if topo == "circular" else n / 8.0)
        g = rng.uniform(0.1, 2.0, 3) * (1.0 if topo == "circular" else n / 2.0)
        prob = qw.Problem(n=n, topology=topo, schedule="cerf_3", n_steps=64, flags=flags)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, topo, "cerf_3", T, g, n_steps=64,
                                             return_psi=True)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
        assert np.abs(psi - psio).max() <= PSI_TOL


def test_gamma_on_laplacian_and_custom_target(qw):
    n = 64
    T, g = np.array([20.0, 35.0]), np.array([1.1 / n, 0.9 / n])
    prob = qw.Problem(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian",
                      n_steps=400)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(n, "complete", "cerf_3", T, g, n_steps=400,
                                   gamma_on="laplacian")
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    prob = qw.Problem(n=40, topology="circular", schedule="lin", target=3, n_steps=200)
    p, norm, psi = qw.rk4_batch(prob, [9.0], [1.3], return_psi=True)
    po, no, psio, _ = c_oracle.rk4_batch(40, "circular", "lin", [9.0], [1.3], n_steps=200,
                                         target=3, return_psi=True)
    assert abs(p[0] - po[0]) <= P_TOL and np.abs(psi - psio).max() <= PSI_TOL


def test_heatmap_layout_argmax_and_row_shards(qw):
    """config 2 shape scaled down: cycle N=64, [beta][T] layout, identical argmax."""
    n, S = 64, 256
    t, b = np.linspace(0.1, 64, 20), np.linspace(0, 2.5, 16)
    for sched in ("lin", "cerf_3", "static"):
        prob = qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S)
        p, norm = qw.rk4_heatmap(prob, t, b)
        po, no = c_oracle.heatmap(n, "circular", sched, t, b, n_steps=S)
        assert p.shape == (16, 20)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
        assert np.argmax(p) == np.argmax(po)                       # bit-exact indices
        assert (np.argmax(p, axis=0) == np.argmax(po, axis=0)).all()
        top2 = np.sort(po.ravel())[-2:]
        assert top2[1] - top2[0] > 100 * P_TOL                     # margin makes it meaningful
        lo, _ = qw.rk4_heatmap(prob, t, b, rows=(3, 9))
        assert np.array_equal(lo, p[3:9])                          # shards are bitwise equal
        ph, nh = qw.rk4_heatmap_host(prob, t, b, rows=(3, 9))
        assert np.array_equal(ph, p[3:9]) and np.array_equal(nh, norm[3:9])


def test_step_size_rule_ragged(qw):
    T = np.array([0.0, 0.004, 1.0, 2.5001, 7.0, 0.3])
    g = np.array([1.0, 0.5, 2.0, 0.1, 1.5, 0.7])
    for n, topo in ((9, "circular"), (128, "circular"), (24, "complete")):
        gg = g if topo == "circular" else g * 5
        prob = qw.Problem(n=n, topology=topo, schedule="sqrt", step_size=1e-2)
        p, norm = qw.rk4_batch(prob, T, gg)
        po, no, _ = c_oracle.rk4_batch(n, topo, "sqrt", T, gg, step_size=1e-2)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
        S = np.array([int(x / 1e-2) for x in T])
        order = qw.longest_first_order(S)
        # a handful of host points on a tiny graph take the parallel-in-time kernels, whose sums
        # are regrouped: equal to rounding; with QW_FLAG_FIXED_GEOMETRY (32768) every call form
        # runs the same kernel and the results are bitwise equal
        p2, _ = qw.rk4_batch(prob, T, gg, order=order)
        assert np.abs(p - p2).max() <= 1e-13
        ph, nh = qw.rk4_batch_host(prob, T, gg)
        assert np.array_equal(p, ph) and np.array_equal(norm, nh)
        fixed = qw.Problem(n=n, topology=topo, schedule="sqrt", step_size=1e-2, flags=32768)
        pf, nf = qw.rk4_batch(fixed, T, gg)
        pf2, _ = qw.rk4_batch(fixed, T, gg, order=order)
        pfh, nfh = qw.rk4_batch_host(fixed, T, gg)
        assert np.array_equal(pf, pf2) and np.array_equal(pf, pfh) and np.array_equal(nf, nfh)
        assert np.abs(pf - po).max() <= P_TOL
    prob = qw.Problem(n=16, topology="circular", schedule="lin")
    S = np.array([10, 0, 33, 64, 7, 128])
    p, norm = qw.rk4_batch(prob, T, g, n_steps=S)
    po, no, _ = c_oracle.rk4_batch(16, "circular", "lin", T, g, n_steps=S)
    assert np.abs(p - po).max() <= P_TOL


def test_full_size_configs_on_a_sample(qw):
    """BASELINE sizes, a few points each, against the C oracle."""
    # config 5: cycle N=1024, S=4096 (points taken from the 1024x1024 grid)
    t, b = np.linspace(0.1, 1024, 1024), np.linspace(0, 2.5, 1024)
    T, g = t[[5, 300, 1023]], b[[700, 410, 9]]
    prob = qw.Problem(n=1024, topology="circular", schedule="lin", n_steps=4096)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(1024, "circular", "lin", T, g, n_steps=4096)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    # config 3: complete N=4096, gamma on the Laplacian, S=2048
    n = 4096
    T, g = np.array([0.1, 200.0, 512.0]), np.array([0.5, 1.0, 2.5]) / n
    prob = qw.Problem(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian",
                      n_steps=2048)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(n, "complete", "cerf_3", T, g, n_steps=2048,
                                   gamma_on="laplacian")
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    # config 4: cycle N=256, S=512, noisy (T, gamma) ensemble
    rng = np.random.default_rng(20200912)
    T0, g0 = (np.pi / 2) * 16.0, 1.0
    T = np.clip(T0 * (1 + 0.05 * rng.standard_normal(64)), 0, None)
    g = np.clip(g0 * (1 + 0.05 * rng.standard_normal(64)), 0, None)
    prob = qw.Problem(n=256, topology="circular", schedule="lin", n_steps=512)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(256, "circular", "lin", T, g, n_steps=512)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL


def test_size_independent_properties_at_full_size(qw):
    """Properties that need no oracle: complete graph == 2-variable ODE, mirror symmetry
    of the ring about the marked vertex, norm defect of RK4, p(T=0) = 1/N."""
    n, S = 4096, 2048
    for go, T, g in (("laplacian", 300.0, 1.0 / n), ("oracle", 0.3, 0.8 * n)):
        prob = qw.Problem(n=n, topology="complete", schedule="cerf_3", gamma_on=go, n_steps=S)
        p, norm, psi = qw.rk4_batch(prob, [T], [g], return_psi=True)
        p2, n2 = rk.complete_two_variable(n, "cerf_3", T, g, S, gamma_on=go)
        assert abs(p[0] - p2) <= P_TOL and abs(norm[0] - n2) <= P_TOL
        others = np.delete(psi[0], default_target(n))
        assert np.abs(others - others[0]).max() <= 1e-14
    n = 1024
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=4096)
    p, norm, psi = qw.rk4_batch(prob, [700.0, 0.0], [1.2, 1.2], return_psi=True)
    w, k = default_target(n), np.arange(1, n // 2)
    assert np.abs(psi[0][(w + k) % n] - psi[0][(w - k) % n]).max() <= 1e-13
    assert abs(norm[0] - 1.0) < 1e-4 and 0.0 <= p[0] <= 1.0      # RK4 norm defect at h = 0.17
    assert p[1] == pytest.approx(1.0 / n, abs=1e-15)


def test_rhs_is_linear(qw):
    rng = np.random.default_rng(7)
    n = 300
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for topo in ("circular", "complete"):
        prob = qw.Problem(n=n, topology=topo, schedule="cerf_5", n_steps=1)
        f = lambda v: qw.rhs(prob, 1.3, 4.0, 0.9, v)
        lhs = f(2.5 * x - 1.5j * y)
        assert np.abs(lhs - (2.5 * f(x) - 1.5j * f(y))).max() <= 1e-11 * np.abs(lhs).max()


def test_ensemble_stats_and_robustness(qw, golden):
    meta, arr = golden
    rng = np.random.default_rng(3)
    p = rng.random(100001)
    st = qw.ensemble_stats(p, thresholds=(0.8, 0.9, 0.95, 0.99))
    assert st["mean"] == pytest.approx(p.mean(), abs=1e-13)
    assert st["std"] == pytest.approx(p.std(), abs=1e-13)
    assert st["min"] == p.min() and st["max"] == p.max()
    for th, f in st["fraction_ge"].items():
        assert f == pytest.approx((p >= th).mean(), abs=1e-15)
    prob = arr["rob_prob"]
    for c in meta["robustness"]:
        R, arg = qw.robustness_map(prob, c["variation"])
        assert round(float(R[c["beta_index"], c["time_index"]]), 4) == c["R"]
        for i in range(prob.shape[1]):
            assert arg[i] == rk.first_argmax(prob[:, i])


def test_error_behaviour(qw):
    with pytest.raises(ValueError):
        qw.rk4_batch(qw.Problem(n=2, topology="circular", n_steps=4), [1.0], [1.0])
    with pytest.raises(ValueError):
        qw.rk4_batch(qw.Problem(n=8, topology="star", n_steps=4), [1.0], [1.0])
    with pytest.raises(ValueError):
        qw.rk4_batch(qw.Problem(n=8, schedule="quartic", n_steps=4), [1.0], [1.0])
    with pytest.raises(ValueError):
        qw.rk4_batch(qw.Problem(n=8), [1.0], [1.0])
    p, norm = qw.rk4_batch(qw.Problem(n=8, n_steps=4), np.zeros(0), np.zeros(0))
    assert p.shape == (0,)


def test_plan_reports_native_kernels(qw):
    pl = qw.plan(qw.Problem(n=1024, topology="circular", n_steps=1))
    assert pl["path"] == "horner-cycle" and pl["sites_per_thread"] == 16
    assert pl["threads_per_block"] == 64 and pl["sm_count"] >= 100
    pl = qw.plan(qw.Problem(n=1024, topology="circular", n_steps=1, flags=2048))
    assert pl["path"] == "resident-cycle" and pl["sites_per_thread"] == 16
    pl = qw.plan(qw.Problem(n=1024, topology="circular", n_steps=1, flags=256))
    assert pl["sites_per_thread"] == 8 and pl["threads_per_block"] == 128
    pl = qw.plan(qw.Problem(n=4096, topology="complete", n_steps=1))
    assert pl["path"] == "horner-complete"
    assert pl["threads_per_block"] == 320            # 8 worker warps + scalar warp + table warp
    pl = qw.plan(qw.Problem(n=4096, topology="complete", n_steps=1, flags=128))
    assert pl["threads_per_block"] == 288            # A/B: one auxiliary warp with both duties
    pl = qw.plan(qw.Problem(n=2048, topology="complete", n_steps=1))
    assert pl["threads_per_block"] == 128            # 4 worker warps, duties folded into them
    pl = qw.plan(qw.Problem(n=2048, topology="complete", n_steps=1, flags=128))
    assert pl["threads_per_block"] == 160            # A/B: 4 worker warps + one auxiliary warp
    pl = qw.plan(qw.Problem(n=512, topology="complete", n_steps=1))
    assert pl["threads_per_block"] == 32 and pl["blocks_per_sm"] >= 9
    pl = qw.plan(qw.Problem(n=4096, topology="complete", n_steps=1, flags=2048))
    assert pl["path"] == "resident-complete" and pl["threads_per_block"] == 256
    before = qw.launch_count()
    qw.rk4_batch(qw.Problem(n=16, n_steps=4), [1.0], [1.0])
    assert qw.launch_count() == before + 1


@pytest.mark.parametrize("n", [8, 64, 1000, 1016, 1024, 1040, 2050, 2064, 4100, 8208, 9001])
@pytest.mark.parametrize("kb_flag", [8, 16, 24, 0, 256, 256 | 8, 2048, 2048 | 8, 2048 | 16,
                                     2048 | 24])
def test_streaming_kernel_vs_oracle(qw, n, kb_flag):
    """Tile-fused HBM-streaming kernel (flags: 2 = force it, 8 / 16 / 24 = 1 / 2 / 4 steps per
    pass, default 8 with 16 sites per thread and 4 with 8; 256 = 8 instead of 16 sites per thread
    for n >= 4096; 2048 = stage form instead
    of the nested form, which is used whenever n is a multiple of the sites per thread),
    windows that wrap, ragged last tiles, marked vertex in a halo copy."""
    rng = np.random.default_rng(n + kb_flag)
    T = np.concatenate([[0.0], rng.uniform(1.0, 9.0, 3)])
    g = rng.uniform(0.1, 2.4, 4)
    for sched, S in (("lin", 37), ("cbrt", 24)):
        prob = qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S,
                          flags=2 | kb_flag)
        assert qw.plan(prob)["path"] == "stream"
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, n_steps=S,
                                             return_psi=True)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
        assert np.abs(psi - psio).max() <= PSI_TOL


def test_streaming_kernel_step_size_rule_and_targets(qw):
    T = np.array([0.0, 0.31, 0.117, 0.25])
    g = np.array([1.0, 0.5, 2.0, 0.1])
    for target in (None, 0, 1, 2, 3, 4, 1015, 1016, 2997, 2999):
        prob = qw.Problem(n=3000, topology="circular", schedule="sqrt", step_size=1e-2,
                          target=target, flags=2)
        p, norm = qw.rk4_batch(prob, T, g)
        po, no, _ = c_oracle.rk4_batch(3000, "circular", "sqrt", T, g, step_size=1e-2,
                                       target=target)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL


def test_config5_streaming_variant_on_a_sample(qw):
    """cycle N=65536, n_steps=256 (BASELINE config 5, streaming variant): two grid points
    against the C oracle, plus mirror symmetry about the marked vertex."""
    n, S = 65536, 256
    T, g = np.array([0.1, 64.0]), np.array([0.6, 2.1])
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S)
    assert qw.plan(prob)["path"] == "stream"
    p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
    po, no, _ = c_oracle.rk4_batch(n, "circular", "lin", T, g, n_steps=S)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    w, k = default_target(n), np.arange(1, 4000)
    assert np.abs(psi[1][(w + k) % n] - psi[1][(w - k) % n]).max() <= 1e-13


def test_heatmap_summary(qw):
    rng = np.random.default_rng(11)
    p = rng.random((37, 53))
    p[5, 7] = p[20, 3] = 2.0                      # tie: the first in flattened order wins
    t = np.linspace(0.1, 30, 53)
    s = qw.heatmap_summary(p, t, t_min=12.0)
    assert s["p_max"] == 2.0 and s["argmax"] == (5, 7) and s["iterations"] == 0.5
    tau = np.where(t[None, :] >= 12.0, t[None, :] / p, np.inf)
    j, i = np.unravel_index(np.argmin(tau), tau.shape)
    assert s["tau_argmin"] == (j, i) and s["tau_min"] == tau[j, i]
    s = qw.heatmap_summary(p, t, t_min=1e9)
    assert s["tau_argmin"] == (-1, -1)


def test_ti_exact_golden_expm(qw, golden):
    """Chebyshev propagator against the reference's own scipy expm values."""
    meta, _ = golden
    for c in meta["static"]:
        prob = qw.Problem(n=c["n"], topology="circular", schedule="static")
        p, nrm = qw.ti_exact_batch(prob, [c["time"]], [c["gamma"]])
        assert abs(p[0] - c["p_expm"]) <= 1e-12 and abs(nrm[0] - 1) <= 1e-12, c


@pytest.mark.parametrize("n", [3, 5, 17, 32, 51, 64, 100, 256, 512, 1000, 1024, 2048])
def test_ti_exact_vs_dense_diagonalisation(qw, n):
    rng = np.random.default_rng(n)
    T = np.concatenate([[0.0, 1e-3], rng.uniform(0.5, 90.0, 4), [-7.5]])
    g = np.concatenate([[1.0, 0.3], rng.uniform(0.0, 2.5, 3), [-0.8, 1.7]])
    for go, tgt in (("oracle", None), ("laplacian", None), ("oracle", 1)):
        if tgt is not None and n < 3:
            continue
        prob = qw.Problem(n=n, topology="circular", schedule="static", gamma_on=go, target=tgt)
        p, nrm, psi = qw.ti_exact_batch(prob, T, g, return_psi=True)
        for k in range(len(T)):
            pe, ne, psie = rk.ti_exact(n, T[k], g[k], gamma_on=go, target=tgt)
            assert abs(p[k] - pe) <= 2e-12 and abs(nrm[k] - 1.0) <= 2e-12, (n, go, k)
            assert np.abs(psi[k] - psie).max() <= 2e-12, (n, go, k)
    t, b = np.linspace(0.1, 40, 7), np.linspace(0, 2.5, 5)
    prob = qw.Problem(n=n, topology="circular", schedule="static")
    hp, hn = qw.ti_exact_heatmap(prob, t, b)
    lo, _ = qw.ti_exact_heatmap(prob, t, b, rows=(1, 4))
    assert hp.shape == (5, 7) and np.array_equal(lo, hp[1:4])
    assert abs(hp[3, 2] - rk.ti_exact(n, t[2], b[3])[0]) <= 2e-12


def test_ti_exact_limits(qw):
    prob = qw.Problem(n=64, topology="circular", schedule="static")
    p, nrm = qw.ti_exact_batch(prob, [1200.0, 5000.0], [1.0, 1.0])
    pe, _, _ = rk.ti_exact(64, 1200.0, 1.0)
    assert abs(p[0] - pe) <= 1e-11 and np.isnan(p[1])         # a|t| beyond the Bessel table
    with pytest.raises(ValueError):
        qw.ti_exact_batch(qw.Problem(n=64, topology="complete", schedule="static"), [1.0], [1.0])
    with pytest.raises(ValueError):
        qw.ti_exact_batch(qw.Problem(n=1031, topology="circular", schedule="static"), [1.0], [1.0])


def test_rk45_reproduces_the_reference_as_shipped(qw, golden):
    """Adaptive RK45 with SciPy's controller on the device vs the numbers the reference's own
    evaluate_probability / grid_probability_evaluation printed (tests/golden), and vs SciPy
    over the stencil RHS (same accepted-step sequence: identical nfev)."""
    meta, arr = golden
    for c in meta["rk45"]:
        prob = qw.Problem(n=c["n"], topology=c["topology"], schedule=c["schedule"])
        p, nrm, nfev = qw.rk45_batch(prob, [c["T"]], [c["beta"]], rtol=1e-6, atol=1e-6,
                                     return_nfev=True)
        assert abs(p[0] - c["p"]) <= 1e-9, (c, p[0])
        po, no, _, nf = rk.rk45_scipy(c["n"], c["topology"], c["schedule"], c["T"], c["beta"])
        assert abs(p[0] - po) <= 1e-9 and abs(nrm[0] - no) <= 1e-9 and nfev[0] == nf, c
    hm = [c for c in meta["heatmaps"] if c["kind"] == "RK45"][0]
    t = np.linspace(*hm["time"][:2], int(hm["time"][2]))
    b = np.linspace(*hm["beta"][:2], int(hm["beta"][2]))
    prob = qw.Problem(n=hm["n"], topology="circular", schedule="lin")
    hp, _ = qw.rk45_heatmap(prob, t, b)
    assert np.abs(hp - arr[hm["key"]]).max() <= 1e-9


@pytest.mark.parametrize("n,topo", [(3, "circular"), (16, "circular"), (51, "circular"),
                                    (71, "circular"), (256, "circular"), (1024, "circular"),
                                    (1500, "circular"), (15, "complete"), (64, "complete")])
def test_rk45_vs_scipy_on_the_stencil(qw, n, topo):
    rng = np.random.default_rng(n)
    for sched, tol in (("lin", 1e-6), ("cerf_3", 1e-8), ("sqrt", 1e-6), ("static", 1e-7)):
        T = float(rng.uniform(2.0, 12.0)) / (1.0 if topo == "circular" else n / 4.0)
        g = float(rng.uniform(0.2, 2.4)) * (1.0 if topo == "circular" else n / 3.0)
        prob = qw.Problem(n=n, topology=topo, schedule=sched)
        p, nrm, psi, nfev = qw.rk45_batch(prob, [T, 0.0], [g, g], rtol=tol, atol=tol,
                                          return_psi=True, return_nfev=True)
        po, no, psio, nf = rk.rk45_scipy(n, topo, sched, T, g, rtol=tol, atol=tol)
        assert abs(p[0] - po) <= 1e-9 and np.abs(psi[0] - psio).max() <= 1e-9, (n, topo, sched)
        assert abs(int(nfev[0]) - nf) <= 6, (n, topo, sched, nfev[0], nf)
        assert p[1] == pytest.approx(1.0 / n, abs=1e-15) and nfev[1] == 0


def test_odd_shapes_and_catch_all_paths(qw):
    """Degenerate grids, negative gamma, a complete graph beyond one block (split over blocks;
    the catch-all kernel only when forced), a ring beyond one block with the fixed-step-size rule
    (cluster kernel by default, streaming kernel with ragged steps when forced)."""
    t, b = np.array([3.0]), np.array([-0.7])                    # 1 x 1 heatmap, gamma < 0
    prob = qw.Problem(n=33, topology="circular", schedule="cerf_5", n_steps=50)
    p, nrm = qw.rk4_heatmap(prob, t, b)
    po, no = c_oracle.heatmap(33, "circular", "cerf_5", t, b, n_steps=50)
    assert p.shape == (1, 1) and abs(p[0, 0] - po[0, 0]) <= P_TOL
    p, nrm = qw.rk4_heatmap(prob, np.linspace(0, 4, 7), b)      # one row
    po, no = c_oracle.heatmap(33, "circular", "cerf_5", np.linspace(0, 4, 7), b, n_steps=50)
    assert p.shape == (1, 7) and np.abs(p - po).max() <= P_TOL

    n = 5000                                                    # complete: split / generic kernel
    T, g = np.array([5.0, 30.0]), np.array([1.0, 1.4]) / n
    po, no, _ = c_oracle.rk4_batch(n, "complete", "lin", T, g, n_steps=40, gamma_on="laplacian")
    for flags, path in ((0, "horner-complete"), (1, "generic")):
        prob = qw.Problem(n=n, topology="complete", schedule="lin", gamma_on="laplacian",
                          n_steps=40, flags=flags)
        assert qw.plan(prob)["path"] == path
        p, nrm = qw.rk4_batch(prob, T, g)
        assert np.abs(p - po).max() <= P_TOL and np.abs(nrm - no).max() <= P_TOL

    n = 6000                                                    # ring, step_size: cluster / streaming
    t, b = np.array([0.0, 0.11, 0.5, 0.33]), np.array([0.4, 1.9])
    po, no = c_oracle.heatmap(n, "circular", "sqrt", t, b, step_size=0.02)
    for flags, path in ((0, "cluster-cycle"), (2, "stream")):
        prob = qw.Problem(n=n, topology="circular", schedule="sqrt", step_size=0.02, flags=flags)
        assert qw.plan(prob)["path"] == path
        p, nrm = qw.rk4_heatmap(prob, t, b)
        assert np.abs(p - po).max() <= P_TOL and np.abs(nrm - no).max() <= P_TOL


@pytest.mark.parametrize("n", [17, 33, 51, 71, 100, 200])
def test_any_dimension_packed_kernel_large_and_small_batches(qw, n):
    """Rings that are not R * 2^k in the warp-packed kernel: a batch of 2003 points (ragged last
    warp, T = 0 points, psi) against the oracle, and slices of five points against the same
    points of the big batch (a point's result does not depend on its neighbours in the warp)."""
    P = 2003
    rng = np.random.default_rng(7 * n)
    T = rng.uniform(0.0, 10.0, P)
    g = rng.uniform(0.0, 2.5, P)
    T[::211] = 0.0
    for sched, gamma_on, target, S in (("cbrt", "oracle", None, 70), ("cerf_7", "laplacian", 1, 33)):
        prob = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on,
                          target=target, n_steps=S)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        idx = np.r_[0:24, P - 24:P]
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T[idx], g[idx], n_steps=S,
                                             gamma_on=gamma_on, target=target, return_psi=True)
        assert np.abs(p[idx] - po).max() <= P_TOL and np.abs(norm[idx] - no).max() <= P_TOL
        assert np.abs(psi[idx] - psio).max() <= PSI_TOL
        for lo in range(0, P, 97):
            ps, ns = qw.rk4_batch(prob, T[lo:lo + 5], g[lo:lo + 5])
            assert np.abs(ps - p[lo:lo + 5]).max() <= 1e-13
            assert np.abs(ns - norm[lo:lo + 5]).max() <= 1e-13
