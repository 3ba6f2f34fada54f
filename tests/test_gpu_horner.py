"""GPU parity of the nested-form (Horner) resident kernels (csrc/qw_horner.cu) against the
oracle and against the classic stage-form kernels (QW_FLAG_NO_HORNER = 2048).

Tolerances as everywhere: |dp| <= 1e-10, |dpsi| <= 1e-11, identical argmax."""

import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu

P_TOL = 1e-10
PSI_TOL = 1e-11
NO_HORNER, R8, OCC4, NO_PACK, NO_ROT = 2048, 256, 128, 512, 4096


@pytest.mark.parametrize("n,flags", [(256, NO_PACK), (264, 0), (512, 0), (520, 0), (1024, 0),
                                     (1024, OCC4), (1024, R8), (1536, 0), (1792, 0), (1024, NO_ROT), (1024, R8 | NO_ROT), (2048, NO_ROT), (1040, 0), (2048, 0), (2048, R8),
                                     (3072, 0), (4096, 0), (4096, R8), (4088, 0),
                                     # any dimension: R or R-1 sites per thread, ghost slots
                                     (257, 0), (263, 0), (511, 0), (513, 0), (999, 0), (1001, 0),
                                     (1023, R8), (1025, 0), (2049, 0), (3001, R8), (4095, 0)])
def test_cycle_all_schedules_vs_oracle(qw, n, flags):
    pl = qw.plan(qw.Problem(n=n, topology="circular", step_size=0.1, flags=flags | NO_PACK))
    assert pl["path"] == "horner-cycle", pl
    rng = np.random.default_rng(n + flags)
    for sched in ["lin", "sqrt", "cbrt", "cerf_3", "cerf_5", "cerf_7", "static"]:
        for gamma_on in ("oracle", "laplacian"):
            P = 4
            T = rng.uniform(0.0, 12.0, P)
            g = rng.uniform(0.0, 2.5, P)
            T[0] = 0.0
            g[1] = 0.0                 # gamma = 0: no oracle (or no Laplacian) at all
            S = 100                    # > 3 coefficient-table chunks, ragged tail
            prob = qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S,
                              gamma_on=gamma_on, flags=flags)
            p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
            po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, n_steps=S,
                                                 gamma_on=gamma_on, return_psi=True)
            assert np.abs(p - po).max() <= P_TOL, (n, sched, gamma_on)
            assert np.abs(norm - no).max() <= P_TOL, (n, sched, gamma_on)
            assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched, gamma_on)


def test_targets_step_size_rule_and_long_runs(qw):
    n = 1024
    for target in (0, 1, 2, 3, 4, 511, 1020, 1023):
        prob = qw.Problem(n=n, topology="circular", schedule="lin", target=target, n_steps=64)
        p, norm, psi = qw.rk4_batch(prob, [9.0], [1.3], return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", "lin", [9.0], [1.3], n_steps=64,
                                             target=target, return_psi=True)
        assert abs(p[0] - po[0]) <= P_TOL and np.abs(psi - psio).max() <= PSI_TOL, target
    T = np.array([0.0, 0.05, 3.3, 7.77, 12.0])
    g = np.array([1.0, 0.3, 2.2, 1.1, 0.7])
    prob = qw.Problem(n=512, topology="circular", schedule="cerf_3", step_size=0.1)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(512, "circular", "cerf_3", T, g, step_size=0.1)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    # the C5 step count on a few points: rounding must not accumulate over 4096 steps
    T, g = np.array([1024.0, 333.3, 51.14]), np.array([0.2102, 1.7, 0.9213])
    prob = qw.Problem(n=1024, topology="circular", schedule="lin", n_steps=4096)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(1024, "circular", "lin", T, g, n_steps=4096)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL


def test_heatmap_matches_the_stage_form_kernel_and_keeps_the_argmax(qw):
    t, b = np.linspace(0.1, 300.0, 24), np.linspace(0.0, 2.5, 16)
    res = {}
    for flags in (0, NO_HORNER):
        prob = qw.Problem(n=1024, topology="circular", schedule="lin", n_steps=1200, flags=flags)
        res[flags] = qw.rk4_heatmap(prob, t, b)
    assert np.abs(res[0][0] - res[NO_HORNER][0]).max() <= 1e-12
    assert np.abs(res[0][1] - res[NO_HORNER][1]).max() <= 1e-12
    assert np.argmax(res[0][0]) == np.argmax(res[NO_HORNER][0])
    ho, _ = c_oracle.heatmap(1024, "circular", "lin", t, b, n_steps=1200)
    assert np.abs(res[0][0] - ho).max() <= P_TOL and np.argmax(res[0][0]) == np.argmax(ho)


@pytest.mark.parametrize("n,flags", [(256, 0), (264, 0), (512, 0), (520, 0), (1024, 0), (1024, R8),
                                     (2048, 0), (2048, R8), (3072, 0), (4096, 0), (4096, R8),
                                     (4088, 0), (257, 0), (511, 0), (1001, 0), (1023, R8),
                                     (2049, 0), (4095, 0), (4096, 128), (3000, 128), (3000, 0),
                                     (512, 128), (1001, 128), (1024, 128), (2048, 128), (1536, 0),
                                     (1500, 0), (600, 0), (256, 128), (300, 128), (300, 0), (400, 0)])
def test_complete_all_schedules_vs_oracle(qw, n, flags):
    # flags 128: one auxiliary warp with both duties (default: scalar warp + table warp above
    # 3072 sites, duties folded into the worker warps up to 3072)
    pl = qw.plan(qw.Problem(n=n, topology="complete", n_steps=1, flags=flags))
    assert pl["path"] == "horner-complete", pl
    rng = np.random.default_rng(7 * n + flags)
    for sched in ["lin", "sqrt", "cbrt", "cerf_3", "cerf_5", "cerf_7", "static"]:
        for gamma_on in ("oracle", "laplacian"):
            P = 4
            T = rng.uniform(0.0, 40.0 / n, P) if gamma_on == "oracle" else rng.uniform(0.0, 30.0, P)
            g = rng.uniform(0.0, 2.0 * n, P) if gamma_on == "oracle" else rng.uniform(0.2, 1.6, P) / n
            T[0] = 0.0
            g[1] = 0.0
            S = 100
            prob = qw.Problem(n=n, topology="complete", schedule=sched, n_steps=S,
                              gamma_on=gamma_on, flags=flags)
            p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
            po, no, psio, _ = c_oracle.rk4_batch(n, "complete", sched, T, g, n_steps=S,
                                                 gamma_on=gamma_on, return_psi=True)
            assert np.abs(p - po).max() <= P_TOL, (n, sched, gamma_on)
            assert np.abs(norm - no).max() <= P_TOL, (n, sched, gamma_on)
            assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched, gamma_on)


def test_complete_config3_sample_and_targets(qw):
    n = 4096
    T = np.array([512.0, 333.3, 102.5, 20.0])
    g = np.array([0.0002446, 0.0004897, 0.0005108, 0.0003])
    for sched in ("cerf_3", "lin"):
        prob = qw.Problem(n=n, topology="complete", schedule=sched, gamma_on="laplacian",
                          n_steps=2048)
        p, norm = qw.rk4_batch(prob, T, g)
        po, no, _ = c_oracle.rk4_batch(n, "complete", sched, T, g, n_steps=2048,
                                       gamma_on="laplacian")
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    for target in (0, 5, 1000, 1023):
        prob = qw.Problem(n=1024, topology="complete", schedule="lin", target=target, n_steps=50)
        p, norm, psi = qw.rk4_batch(prob, [0.03], [700.0], return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(1024, "complete", "lin", [0.03], [700.0], n_steps=50,
                                             target=target, return_psi=True)
        assert abs(p[0] - po[0]) <= P_TOL and np.abs(psi - psio).max() <= PSI_TOL, target


@pytest.mark.parametrize("topo,n", [("circular", 1024), ("complete", 1024), ("complete", 4096),
                                    ("circular", 65536)])
def test_chunk_edges_per_point_steps_and_order(qw, topo, n):
    """Step counts around the 32-step coefficient-table chunks (and the 3-deep table ring of
    the complete-graph kernel), per-point n_steps arrays, a block order permutation and the
    host-buffer C entry point."""
    S = np.array([0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 129], dtype=np.int64)
    P = S.size
    rng = np.random.default_rng(n)
    if topo == "circular":                 # keep every step inside RK4's stability region
        h, g = rng.uniform(0.05, 0.25, P), rng.uniform(0.0, 2.5, P)
    else:
        h, g = rng.uniform(0.2, 1.0, P) / n, rng.uniform(0.0, 2.0 * n, P)
    T = np.where(S > 0, h * S, 1.0)
    prob = qw.Problem(n=n, topology=topo, schedule="cerf_5")
    order = rng.permutation(P)
    p, norm, psi = qw.rk4_batch(prob, T, g, n_steps=S, order=order, return_psi=True)
    po, no, psio, _ = c_oracle.rk4_batch(n, topo, "cerf_5", T, g, n_steps=S, return_psi=True)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    assert np.abs(psi - psio).max() <= PSI_TOL
    assert p[0] == pytest.approx(1.0 / n, abs=1e-15)
    ph, nh = qw.rk4_batch_host(prob, T, g, n_steps=S)
    assert np.array_equal(ph, p) and np.array_equal(nh, norm)


def test_randomised_nested_vs_stage_form(qw):
    """Differential test: the nested-form kernels against the stage-form kernels
    (QW_FLAG_NO_HORNER) on random problems -- dimension, graph, schedule, gamma placement, step
    rule, marked vertex -- at 1e-12, far below the parity tolerance."""
    rng = np.random.default_rng(20240607)
    scheds = ["lin", "sqrt", "cbrt", "cerf_3", "cerf_5", "cerf_7", "static"]
    for trial in range(40):
        topo = "circular" if trial % 2 == 0 else "complete"
        n = int(rng.choice([256, 264, 301, 512, 777, 1024, 1536, 2047, 2048, 4096]))
        sched = scheds[int(rng.integers(len(scheds)))]
        gamma_on = "oracle" if rng.random() < 0.5 else "laplacian"
        target = int(rng.integers(n))
        P = 3
        if topo == "circular":
            h = rng.uniform(0.02, 0.25, P)      # h * 4 * gamma <= 2.5: inside RK4 stability
            g = rng.uniform(0.0, 2.5, P)
        else:
            h = rng.uniform(0.1, 1.0, P) / n
            g = rng.uniform(0.0, 2.0, P) * (n if gamma_on == "oracle" else 1.0 / n)
            if gamma_on == "laplacian":
                h = h * n * rng.uniform(0.3, 1.0)       # a = (1-g) gamma ~ 1/n: h up to ~1
        steps = rng.integers(1, 200, P)
        if rng.random() < 0.5:
            kw, T = dict(n_steps=None), h * steps
            extra = dict(n_steps=steps.astype(np.int64))
        else:
            hs = float(h[0])
            kw, T = dict(step_size=hs), hs * steps + rng.uniform(0.0, 0.9, P) * hs
            extra = {}
        res = []
        for flags in (0, NO_HORNER):
            prob = qw.Problem(n=n, topology=topo, schedule=sched, gamma_on=gamma_on,
                              target=target, flags=flags, **{k: v for k, v in kw.items() if v})
            res.append(qw.rk4_batch(prob, T, g, return_psi=True, **extra))
        ctx = (trial, topo, n, sched, gamma_on, target)
        assert np.abs(res[0][0] - res[1][0]).max() <= 1e-12, ctx
        assert np.abs(res[0][1] - res[1][1]).max() <= 1e-12, ctx
        assert np.abs(res[0][2] - res[1][2]).max() <= 1e-12, ctx


@pytest.mark.parametrize("n", [128, 256, 512])
def test_packed_nested_form_vs_oracle(qw, n):
    """Warp-packed nested-form kernel (N = 128 / 256 / 512, batches that fill the machine): all
    schedules, both gamma placements, T = 0 and gamma = 0 points, ragged last warp, chunk edges
    of the per-point tables (8 / 16 steps), psi output, custom target."""
    P = 1303
    assert qw.plan(qw.Problem(n=n, n_steps=8))["path"] == "packed-horner-cycle"
    rng = np.random.default_rng(n)
    for sched, gamma_on, S, target in [("lin", "oracle", 37, None), ("cerf_3", "laplacian", 16, 5),
                                       ("sqrt", "oracle", 17, 0), ("cbrt", "laplacian", 33, n - 1),
                                       ("cerf_7", "oracle", 8, 2), ("static", "oracle", 9, None),
                                       ("cerf_5", "laplacian", 1, 77)]:
        T = rng.uniform(0.1, 0.2 * S, P)
        g = rng.uniform(0.0, 2.5, P)
        T[::97] = 0.0
        g[5::131] = 0.0
        prob = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on,
                          target=target, n_steps=S)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, n_steps=S,
                                             gamma_on=gamma_on, target=target, return_psi=True)
        assert np.abs(p - po).max() <= P_TOL, (n, sched)
        assert np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched)
    # the same batch through the stage-form packed kernel
    prob2 = qw.Problem(n=n, topology="circular", schedule="cerf_5", gamma_on="laplacian",
                       target=77, n_steps=1, flags=NO_HORNER)
    p2, _ = qw.rk4_batch(prob2, T, g)
    assert np.abs(p2 - p).max() <= 1e-13


def test_runs_are_bitwise_reproducible(qw):
    """Every nested-form kernel twice on the same inputs: bitwise identical p, norm and psi
    (fixed summation orders; a shared-memory race would show up here as run-to-run noise)."""
    rng = np.random.default_rng(99)
    cases = [dict(n=1024, topology="circular", n_steps=200),
             dict(n=2048, topology="circular", n_steps=90),
             dict(n=4096, topology="circular", n_steps=70),
             dict(n=1024, topology="complete", n_steps=200),
             dict(n=4096, topology="complete", n_steps=130),
             dict(n=256, topology="circular", n_steps=100),
             dict(n=8192, topology="circular", n_steps=24)]
    for kw in cases:
        P = 700 if kw["n"] == 256 else 40
        scale = 1.0 if kw["topology"] == "circular" else 1.0 / kw["n"]
        T = rng.uniform(1.0, 9.0, P) * scale
        g = rng.uniform(0.1, 2.4, P) / scale
        prob = qw.Problem(schedule="cerf_3", **kw)
        a = qw.rk4_batch(prob, T, g, return_psi=True)
        for _ in range(3):
            b = qw.rk4_batch(prob, T, g, return_psi=True)
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), kw


@pytest.mark.parametrize("n", [5, 16, 51, 64, 71, 100, 128, 255, 256, 512])
def test_packed_kernels_fixed_step_size_on_heatmaps(qw, n):
    """Fixed step size (the drop-in modules' default rule, QW_Search.py:153-154) on a heatmap:
    the packed kernels take the points of one T column per warp (common step count and stage
    times).  Ascending and descending T, T = 0, row shards, a batch large enough for the nested
    packed kernel at N = 128 / 256 / 512."""
    rng = np.random.default_rng(n)
    nT, nB = 9, (160 if n >= 128 else 23)
    t = np.sort(rng.uniform(0.3, 6.0, nT))
    t[0] = 0.0
    b = rng.uniform(0.0, 2.5, nB)
    for sched, order in (("lin", 1), ("cbrt", -1)):
        tt = t[::order].copy()
        prob = qw.Problem(n=n, topology="circular", schedule=sched, step_size=0.05)
        assert qw.plan(prob)["path"].startswith("packed")
        p, norm = qw.rk4_heatmap(prob, tt, b)
        po, no = c_oracle.heatmap(n, "circular", sched, tt, b, step_size=0.05)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.argmax(p) == np.argmax(po)
        p2, _ = qw.rk4_heatmap(prob, tt, b, rows=(3, 17))      # small batch: another geometry
        assert np.abs(p2 - p[3:17]).max() <= 1e-13


MIRROR = 8192


@pytest.mark.parametrize("n", [160, 161, 255, 256, 511, 512, 574, 575, 576, 600, 703, 704, 777, 831,
                               832, 900, 959, 960, 1000, 1023, 1024, 1085, 1086])
def test_mirror_reduced_kernel_vs_oracle(qw, n):
    """QW_FLAG_MIRROR: the half ring with reflecting ends, one warp per point, against the
    full-ring oracle -- even and odd n, every register geometry (9 .. 17 sites per lane), all
    schedules, both gamma placements, both step rules, custom targets, psi (both images)."""
    pl = qw.plan(qw.Problem(n=n, topology="circular", n_steps=1, flags=MIRROR))
    assert pl["path"] == "mirror-cycle", pl
    if n >= 544:        # one point per warp: the fewest sites per lane that fit 32 lanes
        m = n // 2 + 1
        assert pl["steps_per_pass"] == 1, pl
        assert pl["sites_per_thread"] == next(R for R in (9, 11, 13, 15, 17) if -(-m // R) <= 32)
    rng = np.random.default_rng(n)
    for sched, gamma_on, target, rule in [("lin", "oracle", None, dict(n_steps=100)),
                                          ("cbrt", "laplacian", 0, dict(n_steps=37)),
                                          ("cerf_5", "oracle", n - 1, dict(step_size=0.07)),
                                          ("static", "laplacian", 7, dict(n_steps=64))]:
        T = rng.uniform(0.0, 9.0, 5)
        g = rng.uniform(0.0, 2.5, 5)
        T[0], g[1] = 0.0, 0.0
        prob = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on,
                          target=target, flags=MIRROR, **rule)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, gamma_on=gamma_on,
                                             target=target, return_psi=True, **rule)
        assert np.abs(p - po).max() <= P_TOL, (n, sched)
        assert np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched)
    # heatmap through the grid entry point, identical argmax
    t, b = np.linspace(0.1, 40.0, 7), np.linspace(0.0, 2.5, 9)
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=300, flags=MIRROR)
    hp, _ = qw.rk4_heatmap(prob, t, b)
    ho, _ = c_oracle.heatmap(n, "circular", "lin", t, b, n_steps=300)
    assert np.abs(hp - ho).max() <= P_TOL and np.argmax(hp) == np.argmax(ho)


def test_mirror_flag_is_ignored_where_it_does_not_apply(qw):
    for kw in (dict(n=64, topology="circular"), dict(n=159, topology="circular")):
        prob = qw.Problem(n_steps=50, flags=MIRROR, **kw)
        assert qw.plan(prob)["path"] not in ("mirror-cycle", "symmetric-complete")
        scale = 1.0 if kw["topology"] == "circular" else 1.0 / kw["n"]
        p, norm = qw.rk4_batch(prob, [3.0 * scale], [1.2 / scale])
        po, no, _ = c_oracle.rk4_batch(kw["n"], kw["topology"], "lin", [3.0 * scale],
                                       [1.2 / scale], n_steps=50)
        assert abs(p[0] - po[0]) <= P_TOL


def test_localization_run_131072_steps(qw):
    """BASELINE config 4, localization ensemble: T ~ N^2/2 = 32768, 131072 steps.  Rounding must
    not accumulate in the nested form, the stage form or the mirror reduction."""
    rng = np.random.default_rng(5)
    P = 700                                        # fills the machine: packed nested kernel
    T = 32768.0 * (1 + 0.05 * rng.standard_normal(P))
    g = 1.45 * (1 + 0.05 * rng.standard_normal(P))
    po, no, _ = c_oracle.rk4_batch(256, "circular", "lin", T[:4], g[:4], n_steps=131072)
    for flags in (0, NO_HORNER, MIRROR):
        prob = qw.Problem(n=256, topology="circular", schedule="lin", n_steps=131072, flags=flags)
        p, nrm = qw.rk4_batch(prob, T, g)
        assert np.abs(p[:4] - po).max() <= 1e-11 and np.abs(nrm[:4] - no).max() <= 1e-11, flags


@pytest.mark.parametrize("n", [8704, 9001, 10000, 20011])
@pytest.mark.parametrize("kb_flag", [8, 16, 24, 0, 256])
def test_mirror_reduced_streaming_vs_oracle(qw, n, kb_flag):
    """QW_FLAG_MIRROR in the HBM-streaming kernel: the half ring behind three ghost sites,
    reflecting window halos; even and odd n, 1 / 2 / 4 / 8 steps per pass, both window geometries,
    custom target, psi with both mirror images."""
    rng = np.random.default_rng(n + kb_flag)
    T = np.concatenate([[0.0], rng.uniform(1.0, 9.0, 3)])
    g = rng.uniform(0.1, 2.4, 4)
    for sched, S, target in (("lin", 37, None), ("cbrt", 24, 5)):
        prob = qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S, target=target,
                          flags=MIRROR | kb_flag)
        assert qw.plan(prob)["path"] == "stream"
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, n_steps=S, target=target,
                                             return_psi=True)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched)


FIXED, NO_PACK, R8 = 32768, 512, 256


def _mirror_packed_geometry(n, r8=False):
    """(sites per lane, points per warp) as mirror_packed_geometry of qw_packed.cu picks them."""
    m = n // 2 + 1
    r1 = next(R for R in (9, 11, 13, 15, 17) if -(-m // R) <= 32)     # one point per warp
    best, best_score = (r1, 1), m * 110 // r1
    for R, lp in ((17, 8), (15, 8), (13, 8), (17, 16), (15, 16), (13, 16), (9, 16)):
        L = -(-m // R)
        if (R != 9 and r8) or L > lp or m < L * (R - 1) + 1:
            continue
        score = (32 // lp) * m * (100 + (R - 9) * 10 // 8) // R
        if score > best_score:
            best, best_score = (R, 32 // lp), score
    return best


def test_mirror_packed_geometry_examples():
    assert [_mirror_packed_geometry(n) for n in (160, 178, 200, 256, 271, 286, 288, 310, 350,
                                                 512, 542, 544, 576)] == \
        [(17, 4), (13, 4), (13, 4), (17, 4), (17, 4), (9, 2), (13, 2), (13, 2), (13, 2),
         (17, 2), (17, 2), (9, 1), (11, 1)]
    assert _mirror_packed_geometry(256, r8=True) == (9, 2)


@pytest.mark.parametrize("n,flags", [(160, 0), (161, 0), (170, 0), (200, 0), (200, R8), (255, 0),
                                     (256, 0), (256, R8), (257, 0), (270, 0), (271, 0), (272, 0),
                                     (286, 0), (287, 0), (288, 0), (300, 0), (305, 0), (310, 0),
                                     (178, 0), (179, 0), (172, 0), (188, 0), (220, 0), (238, 0),
                                     (318, 0), (350, 0), (383, 0), (415, 0), (447, 0), (478, 0),
                                     (480, 0), (512, 0), (513, 0), (542, 0), (543, 0), (574, 0)])
def test_mirror_reduced_kernel_points_packed_per_warp(qw, n, flags):
    """Half rings of at most 16 lanes: two or four points per warp (BASELINE config 4 with the
    drop-in default).  QW_FLAG_FIXED_GEOMETRY gives small calls the geometry of large ones, so
    the packed kernel is what runs here: against the full-ring oracle and against the
    one-point-per-warp kernel (QW_FLAG_NO_PACK), with points of T = 0 inside a warp, batch sizes
    that leave groups of the last warp idle, both step rules, targets, psi."""
    r, per_warp = _mirror_packed_geometry(n, bool(flags & R8))
    pl = qw.plan(qw.Problem(n=n, topology="circular", n_steps=1, flags=MIRROR | flags))
    assert pl["path"] == "mirror-cycle" and pl["steps_per_pass"] == per_warp, pl
    assert pl["sites_per_thread"] == r and pl["blocks_per_sm"] == 12, pl
    pl1 = qw.plan(qw.Problem(n=n, topology="circular", n_steps=1, flags=MIRROR | NO_PACK))
    assert pl1["path"] == "mirror-cycle" and pl1["steps_per_pass"] == 1, pl1
    rng = np.random.default_rng(n)
    for sched, gamma_on, target, S, npts in [("lin", "oracle", None, 100, 11),
                                             ("cbrt", "laplacian", 0, 37, 5),
                                             ("cerf_5", "oracle", n - 1, 64, 8),
                                             ("static", "laplacian", 7, 3, 1)]:
        T = rng.uniform(0.0, 9.0, npts)
        g = rng.uniform(0.0, 2.5, npts)
        T[npts // 2] = 0.0
        g[0] = 0.0
        prob = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on,
                          target=target, n_steps=S, flags=MIRROR | FIXED | flags)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, gamma_on=gamma_on,
                                             target=target, return_psi=True, n_steps=S)
        assert np.abs(p - po).max() <= P_TOL, (n, sched)
        assert np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched)
        prob1 = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on,
                           target=target, n_steps=S, flags=MIRROR | NO_PACK)
        p1, norm1 = qw.rk4_batch(prob1, T, g)
        assert np.abs(p - p1).max() <= 1e-13 and np.abs(norm - norm1).max() <= 1e-13
    # heatmaps through the grid entry point: the n_steps rule and a fixed step size (a warp takes
    # rows of one T column), identical argmax, row shards bitwise the rows of the whole
    t, b = np.linspace(0.0, 12.0, 5), np.linspace(0.0, 2.5, 11)
    for rule in (dict(n_steps=48), dict(step_size=0.08)):
        prob = qw.Problem(n=n, topology="circular", schedule="cerf_3", flags=MIRROR | FIXED | flags,
                          **rule)
        p, norm = qw.rk4_heatmap(prob, t, b)
        po, no = c_oracle.heatmap(n, "circular", "cerf_3", t, b, **rule)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL, (n, rule)
        assert np.argmax(p) == np.argmax(po)
        lo, _ = qw.rk4_heatmap(prob, t, b, rows=(3, 8))
        assert np.array_equal(lo, p[3:8])


def _block_mirror_sites_per_thread(n):
    """qw_horner_sites_per_thread under QW_FLAG_MIRROR: filled slots x weight of R x weight of
    the warps resident per SM, ties to the larger R."""
    m = n // 2 + 1
    score = {}
    for R, w in ((17, 106), (13, 105), (9, 95)):
        nact = -(-m // R)
        if nact > 256:
            continue
        warps = -(-nact // 32)
        blocks = (4 if R == 17 else 6) if warps <= 2 else (2 if R == 17 else 3) if warps <= 4 else 1
        resident = blocks * warps
        occ = 100 if resident >= 8 else {7: 92, 6: 80}.get(resident, 60)
        score[R] = m * w * occ * (90 if warps == 3 else 100) // (R * warps * 32)
    return max(score, key=lambda R: (score[R], R))


def test_block_mirror_geometry_examples():
    assert [_block_mirror_sites_per_thread(n) for n in (1088, 1100, 1300, 2048, 2200, 2500, 3000,
                                                        3300, 4096, 4606, 5000, 6656, 8192,
                                                        8702)] == \
        [9, 9, 13, 17, 9, 13, 13, 13, 17, 9, 13, 17, 17, 17]


@pytest.mark.parametrize("n,flags", [(1087, 0), (1088, 0), (1100, 0), (1300, 0), (1664, 0),
                                     (2048, 0), (2200, 0), (2500, 0), (3000, 0), (3300, 0),
                                     (4096, 0), (4097, 0), (4606, 0), (4607, 0), (5000, 0), (6656, 0),
                                     (8190, 0), (8191, 0), (8702, 0), (2200, 512)])
def test_mirror_reduced_block_kernel_vs_oracle(qw, n, flags):
    """QW_FLAG_MIRROR between 1087 and 8191 sites: the half ring in the block-per-point nested
    kernel (reflecting ends in thread 0 and in the last thread)."""
    pl = qw.plan(qw.Problem(n=n, topology="circular", n_steps=1, flags=MIRROR | flags))
    assert pl["path"] == ("mirror-cycle" if n // 2 + 1 <= 544 else "horner-cycle"), pl
    if n // 2 + 1 > 544:
        assert pl["sites_per_thread"] == (17 if flags & 512 else _block_mirror_sites_per_thread(n))
    rng = np.random.default_rng(n)
    for sched, gamma_on, target, rule in [("lin", "oracle", None, dict(n_steps=70)),
                                          ("sqrt", "laplacian", 3, dict(n_steps=33)),
                                          ("cerf_3", "oracle", n - 2, dict(step_size=0.11))]:
        T = rng.uniform(0.0, 7.0, 4)
        g = rng.uniform(0.0, 2.5, 4)
        T[0], g[1] = 0.0, 0.0
        prob = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on,
                          target=target, flags=MIRROR | flags, **rule)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, gamma_on=gamma_on,
                                             target=target, return_psi=True, **rule)
        assert np.abs(p - po).max() <= P_TOL, (n, sched)
        assert np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 256, 1000, 4096, 5000])
def test_symmetry_reduced_complete_graph_vs_oracle(qw, n):
    """QW_FLAG_MIRROR on the complete graph: one thread integrates (psi_w, psi_other) of a point.
    Against the full-vector oracle: all schedules, both gamma placements, both step rules, custom
    targets, psi expanded to all vertices; then against the default block-per-point kernel on a
    heatmap (identical argmax)."""
    pl = qw.plan(qw.Problem(n=n, topology="complete", n_steps=1, flags=MIRROR))
    assert pl["path"] == "symmetric-complete", pl
    rng = np.random.default_rng(n)
    for sched, gamma_on, target, rule in [("lin", "oracle", None, dict(n_steps=100)),
                                          ("cbrt", "laplacian", 0, dict(n_steps=37)),
                                          ("cerf_5", "oracle", n - 1, dict(step_size=0.07 / n)),
                                          ("sqrt", "oracle", n // 3, dict(n_steps=45)),
                                          ("static", "laplacian", n // 2, dict(n_steps=64))]:
        T = rng.uniform(0.0, 9.0, 37) / n                 # h * N stays in the stable range
        g = rng.uniform(0.0, 2.5, 37) * (n if gamma_on == "oracle" else 1.0 / n)
        T[0], g[1] = 0.0, 0.0
        prob = qw.Problem(n=n, topology="complete", schedule=sched, gamma_on=gamma_on,
                          target=target, flags=MIRROR, **rule)
        p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "complete", sched, T, g, gamma_on=gamma_on,
                                             target=target, return_psi=True, **rule)
        assert np.abs(p - po).max() <= P_TOL, (n, sched)
        assert np.abs(norm - no).max() <= P_TOL, (n, sched)
        assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched)
    if n >= 256:
        t, b = np.linspace(0.1, 80.0, 33) / n, np.linspace(0.0, 2.5, 41) * n
        for rule in (dict(n_steps=400), dict(step_size=0.2 / n)):
            full, nf = qw.rk4_heatmap(qw.Problem(n=n, topology="complete", schedule="lin", **rule),
                                      t, b)
            red, nr = qw.rk4_heatmap(qw.Problem(n=n, topology="complete", schedule="lin",
                                                flags=MIRROR, **rule), t, b)
            assert np.abs(full - red).max() <= 1e-12 and np.abs(nf - nr).max() <= 1e-12
            assert np.argmax(full) == np.argmax(red)


def test_symmetry_reduced_complete_graph_beyond_memory(qw):
    """With the reduction the dimension is just a number: N = 10^9 vertices (the full vector
    would be 16 GB per point).  Checked against the same 2x2 system integrated with the
    reference's RK4 in numpy, relative to the size of p (which is O(1/N) here)."""
    n, S = 10 ** 9, 4000
    T, gam = np.array([3000.0, 5000.0]), np.array([1.0, 1.3])
    prob = qw.Problem(n=n, topology="complete", schedule="lin", gamma_on="laplacian", n_steps=S,
                      flags=MIRROR)
    p, norm = qw.rk4_batch(prob, T, gam / n)
    for q in range(2):
        m, h = n - 1.0, T[q] / S
        y = np.array([1.0, 1.0], dtype=complex) / np.sqrt(n)

        def rhs(t, y):
            s = t / T[q]
            a, d = (1.0 - s) * gam[q] / n, -s
            dlt = y[0] - y[1]
            return -1j * np.array([a * m * dlt + d * y[0], -a * dlt])
        for k in range(S):
            t = k * h
            k1 = rhs(t, y)
            k2 = rhs(t + 0.5 * h, y + 0.5 * h * k1)
            k3 = rhs(t + 0.5 * h, y + 0.5 * h * k2)
            k4 = rhs(t + h, y + h * k3)
            y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        nrm = abs(y[0]) ** 2 + m * abs(y[1]) ** 2
        pref = abs(y[0]) ** 2 / nrm
        assert abs(p[q] - pref) <= 1e-9 * pref and abs(norm[q] - nrm) <= 1e-10


@pytest.mark.parametrize("n", [256, 300, 512, 600, 1000, 1024, 2048, 2560, 3000, 4096])
def test_complete_kernels_repeatable_and_variants_agree(qw, n):
    """The chunk-ahead scheme of the complete-graph kernels (tables two chunks ahead, scalar steps
    one chunk ahead, one barrier per chunk; folded into the worker warps up to 3072 sites, one
    map buffer in one-warp blocks) has no ordering left to chance: a full machine of blocks gives
    bitwise the same heatmap every time, and the one-auxiliary-warp variant (flags 128) agrees to
    rounding.  235 steps = 7 full chunks and a ragged one."""
    import torch

    t = torch.linspace(0.5, 40.0, 96, dtype=torch.float64, device="cuda")
    b = torch.linspace(0.3 / n, 2.2 / n, 32, dtype=torch.float64, device="cuda")
    out = {}
    for flags in (0, 128):
        prob = qw.Problem(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian",
                          n_steps=235, flags=flags)
        assert qw.plan(prob)["path"] == "horner-complete"
        runs = [qw.rk4_heatmap(prob, t, b) for _ in range(4)]
        p0, n0 = runs[0][0].cpu().numpy(), runs[0][1].cpu().numpy()
        for p, nrm in runs[1:]:
            assert np.array_equal(p.cpu().numpy(), p0) and np.array_equal(nrm.cpu().numpy(), n0)
        out[flags] = (p0, n0)
    assert np.abs(out[0][0] - out[128][0]).max() <= 1e-13
    assert np.abs(out[0][1] - out[128][1]).max() <= 1e-13
    sel = [(0, 0), (31, 95), (17, 40)]
    po, no, _ = c_oracle.rk4_batch(n, "complete", "cerf_3", [float(t[j]) for _, j in sel],
                                   [float(b[i]) for i, _ in sel], n_steps=235,
                                   gamma_on="laplacian")
    for k, (i, j) in enumerate(sel):
        assert abs(out[0][0][i, j] - po[k]) <= P_TOL and abs(out[0][1][i, j] - no[k]) <= P_TOL


@pytest.mark.parametrize("n", [4, 5, 15, 16, 17, 51, 64, 71, 100, 128, 200, 255])
def test_packed_complete_small_graphs_vs_oracle(qw, n):
    """Warp-packed kernel for small complete graphs (4 <= n < 256; the sizes the reference runs):
    R or R-1 sites per lane, several points per warp, stage sums tracked through the marked
    vertex.  All schedules, both gamma placements, psi, ragged batches, custom targets."""
    pl = qw.plan(qw.Problem(n=n, topology="complete", n_steps=1))
    assert pl["path"] == "packed-complete", pl
    assert qw.plan(qw.Problem(n=n, topology="complete", n_steps=1, flags=512))["path"] != \
        "packed-complete"
    rng = np.random.default_rng(11 * n)
    for sched in ["lin", "sqrt", "cbrt", "cerf_3", "cerf_5", "cerf_7", "static"]:
        for gamma_on in ("oracle", "laplacian"):
            P = 37                                   # not a multiple of the points per warp
            T = rng.uniform(0.0, 40.0 / n, P) if gamma_on == "oracle" else rng.uniform(0.0, 30.0, P)
            g = rng.uniform(0.0, 2.0 * n, P) if gamma_on == "oracle" else rng.uniform(0.2, 1.6, P) / n
            T[0] = 0.0
            g[1] = 0.0
            S = 100
            prob = qw.Problem(n=n, topology="complete", schedule=sched, n_steps=S,
                              gamma_on=gamma_on)
            p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
            po, no, psio, _ = c_oracle.rk4_batch(n, "complete", sched, T, g, n_steps=S,
                                                 gamma_on=gamma_on, return_psi=True)
            assert np.abs(p - po).max() <= P_TOL, (n, sched, gamma_on)
            assert np.abs(norm - no).max() <= P_TOL, (n, sched, gamma_on)
            assert np.abs(psi - psio).max() <= PSI_TOL, (n, sched, gamma_on)
    for target in (0, n - 1, n // 3):
        prob = qw.Problem(n=n, topology="complete", schedule="lin", target=target, n_steps=60)
        p, norm, psi = qw.rk4_batch(prob, [0.7 / n, 3.0 / n], [0.9 * n, 1.4 * n], return_psi=True)
        po, no, psio, _ = c_oracle.rk4_batch(n, "complete", "lin", [0.7 / n, 3.0 / n],
                                             [0.9 * n, 1.4 * n], n_steps=60, target=target,
                                             return_psi=True)
        assert np.abs(p - po).max() <= P_TOL and np.abs(psi - psio).max() <= PSI_TOL


def test_packed_complete_step_size_heatmap_and_long_run(qw):
    """Fixed step size on a heatmap (a warp takes rows of one T column; the rule the drop-in
    modules default to), and 60 000 steps without re-anchoring the tracked sums."""
    n = 51
    t = np.linspace(0.05, 9.0, 23)
    b = np.linspace(0.2, 1.9, 13) / n
    prob = qw.Problem(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian",
                      step_size=0.01)
    assert qw.plan(prob)["path"] == "packed-complete"
    p, norm = qw.rk4_heatmap(prob, t, b)
    ho, _ = c_oracle.heatmap(n, "complete", "cerf_3", t, b, step_size=0.01, gamma_on="laplacian")
    assert np.abs(p - ho).max() <= P_TOL and np.argmax(p) == np.argmax(ho)
    T = [600.0, 451.0]
    g = [1.0 / n, 0.7 / n]
    prob = qw.Problem(n=n, topology="complete", schedule="lin", gamma_on="laplacian",
                      n_steps=60000)
    pl, nl, psil = qw.rk4_batch(prob, T, g, return_psi=True)
    po, no, psio, _ = c_oracle.rk4_batch(n, "complete", "lin", T, g, n_steps=60000,
                                         gamma_on="laplacian", return_psi=True)
    assert np.abs(pl - po).max() <= 1e-11 and np.abs(nl - no).max() <= 1e-11
    assert np.abs(psil - psio).max() <= 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("n,target", [(128, None), (256, 7), (512, None), (520, 100), (1001, None),
                                      (1024, None), (1024, 0), (2048, 2047), (3000, None),
                                      (4096, 1234), (5000, None), (8192, 3)])
def test_full_ring_state_is_mirror_symmetric_bit_for_bit(qw, n, target):
    """The oracle correction of the nested kernels reads only the sites w-3 .. w (s_q = 2 y_{w-q},
    qw_horner_level.cuh): the full-ring state must be symmetric about the marked vertex BITWISE --
    uniform start (BCB.py:179-180), reflection-invariant H(t), commutative neighbour sums -- in the
    warp-packed, block-per-point and cluster kernels, for ragged layouts and any target."""
    kw = dict(n=n, topology="circular", schedule="cerf_3", n_steps=96)
    if target is not None:
        kw["target"] = target
    prob = qw.Problem(**kw)
    T, g = np.array([3.0, 11.0, 40.0]), np.array([0.4, 1.1, 2.3])
    p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
    w = (n - 1) // 2 if target is None else target
    k = np.arange(1, n // 2 + 1)
    assert np.array_equal(psi[:, (w + k) % n], psi[:, (w - k) % n]), (n, target)
    po, no, psio, _ = c_oracle.rk4_batch(n, "circular", "cerf_3", T, g, n_steps=96,
                                         target=w, return_psi=True)
    assert np.abs(p - po).max() <= P_TOL and np.abs(psi - psio).max() <= PSI_TOL
