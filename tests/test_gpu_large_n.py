"""Dimensions above one thread block's registers (4096 < N <= 32768 and beyond) against the C
oracle: the complete graph split over blocks that share only the tabulated (y_w, S) subsystem,
and the cycle graph resident across a thread-block cluster with DSMEM halos (north star:
"resident ... for N up to about 8K")."""
import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu

P_TOL, PSI_TOL = 1e-10, 1e-11


def _split_blocks(n):
    """qw_horner_split: blocks of 1024 .. 2048 sites, cost = blocks x warps x cost of a warp."""
    ceil = lambda a, b: -(-a // b)
    warp_cost = {2: 106, 3: 118, 4: 100}
    cost = {}
    for s in range(ceil(n, 2048), ceil(n, 1024) + 1):
        warps = ceil(ceil(ceil(n, s), 16), 32)
        cost[s] = s * warps * warp_cost[warps]
    return min(cost, key=lambda s: (cost[s], s))


def test_split_rule_examples():
    assert [_split_blocks(n) for n in (4097, 5000, 6144, 8192, 9000, 12289, 16384, 40000)] == \
        [5, 5, 3, 4, 9, 13, 8, 20]


@pytest.mark.parametrize("n", [4097, 5000, 6144, 8192, 12289, 16384, 40000])
def test_complete_graph_split_over_blocks(qw, n):
    rng = np.random.default_rng(n)
    for sched, go, S in (("cerf_3", "laplacian", 70), ("lin", "oracle", 33), ("sqrt", "oracle", 64)):
        T = np.concatenate([[0.0], rng.uniform(10.0, 60.0, 3)]) if go == "laplacian" else \
            np.concatenate([[0.0], rng.uniform(0.5, 3.0, 3) / n * 8])
        g = rng.uniform(0.5, 2.0, 4) / n if go == "laplacian" else rng.uniform(0.3, 1.5, 4) * n
        po, no, psio, _ = c_oracle.rk4_batch(n, "complete", sched, T, g, n_steps=S, gamma_on=go,
                                             return_psi=True)
        # default: the block count of 1024 .. 2048 sites each that idles the fewest lanes;
        # QW_FLAG_OCC4 (128): blocks of 2049 .. 4096 sites (qw_horner_split)
        for flags, blocks in ((0, _split_blocks(n)), (128, -(-n // 4096))):
            prob = qw.Problem(n=n, topology="complete", schedule=sched, gamma_on=go, n_steps=S,
                              flags=flags)
            pl = qw.plan(prob)
            assert pl["path"] == "horner-complete" and pl["steps_per_pass"] == blocks
            p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
            assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
            assert np.abs(psi - psio).max() <= PSI_TOL


def test_complete_graph_split_step_size_rule_target_and_grid(qw):
    n = 9000
    T = np.array([0.0, 0.0031, 0.00117, 0.0025])
    g = np.array([1.0, 0.5, 2.0, 0.1]) * n
    for target in (None, 0, 4500, 8999):
        prob = qw.Problem(n=n, topology="complete", schedule="sqrt", step_size=1e-4, target=target)
        p, norm = qw.rk4_batch(prob, T, g)
        po, no, _ = c_oracle.rk4_batch(n, "complete", "sqrt", T, g, step_size=1e-4, target=target)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    t, b = np.linspace(0.1, 80, 7), np.linspace(0.5 / n, 2.5 / n, 5)
    prob = qw.Problem(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian", n_steps=96)
    p, norm = qw.rk4_heatmap(prob, t, b)
    po, no = c_oracle.heatmap(n, "complete", "cerf_3", t, b, n_steps=96, gamma_on="laplacian")
    assert np.abs(p - po).max() <= P_TOL and np.argmax(p) == np.argmax(po)
    lo, _ = qw.rk4_heatmap(prob, t, b, rows=(1, 4))
    assert np.array_equal(lo, p[1:4])
    # the two-variable reduction (QW_FLAG_MIRROR) and the catch-all kernel agree with the split one
    ps, _ = qw.rk4_heatmap(qw.Problem(n=n, topology="complete", schedule="cerf_3",
                                      gamma_on="laplacian", n_steps=96, flags=8192), t, b)
    pg, _ = qw.rk4_heatmap(qw.Problem(n=n, topology="complete", schedule="cerf_3",
                                      gamma_on="laplacian", n_steps=96, flags=1), t, b)
    assert np.abs(ps - p).max() <= P_TOL and np.abs(pg - p).max() <= P_TOL


@pytest.mark.parametrize("n", [4097, 5000, 8192, 12289, 16384, 20000, 32768])
def test_cycle_graph_resident_across_a_cluster(qw, n):
    rng = np.random.default_rng(n)
    T = np.concatenate([[0.0], rng.uniform(1.0, 9.0, 4)])
    g = rng.uniform(0.1, 2.4, 5)
    # CTAs per cluster: up to 4096 sites each, or (above 8192 sites, while the cluster stays
    # portable) up to 2050 sites each, two CTAs per SM; 128 = QW_FLAG_OCC4 keeps the big CTAs
    small = -(-n // 2050)
    ctas = {0: small if n > 8192 and small <= 8 else -(-n // 4096), 128: -(-n // 4096)}
    for sched, S in (("lin", 37), ("cbrt", 70)):
        default = qw.plan(qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S))
        assert default["path"] == ("stream" if n > 8192 and n % 128 == 0 else "cluster-cycle")
        po, no, psio, _ = c_oracle.rk4_batch(n, "circular", sched, T, g, n_steps=S,
                                             return_psi=True)
        for geometry in ((0, 128) if ctas[0] != ctas[128] else (0,)):
            # 16384 = QW_FLAG_FORCE_CLUSTER: above 8192 sites rings of a multiple of 128 sites go
            # to the bulk-copy streaming kernel by default
            prob = qw.Problem(n=n, topology="circular", schedule=sched, n_steps=S,
                              flags=16384 | geometry)
            pl = qw.plan(prob)
            assert pl["path"] == "cluster-cycle" and pl["steps_per_pass"] == ctas[geometry]
            p, norm, psi = qw.rk4_batch(prob, T, g, return_psi=True)
            assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
            assert np.abs(psi - psio).max() <= PSI_TOL


def test_cluster_step_size_rule_targets_and_grid(qw):
    n = 9000
    T = np.array([0.0, 0.31, 0.117, 0.25, 0.004])
    g = np.array([1.0, 0.5, 2.0, 0.1, 0.7])
    for target in (None, 0, 1, 2, 3, 4, 4499, 4500, 8997, 8999):
        prob = qw.Problem(n=n, topology="circular", schedule="sqrt", step_size=1e-2, target=target)
        assert qw.plan(prob)["path"] == "cluster-cycle"
        p, norm = qw.rk4_batch(prob, T, g)
        po, no, _ = c_oracle.rk4_batch(n, "circular", "sqrt", T, g, step_size=1e-2, target=target)
        assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
    t, b = np.linspace(0.1, 12, 9), np.linspace(0, 2.5, 7)
    prob = qw.Problem(n=8192, topology="circular", schedule="cerf_3", n_steps=96)
    p, norm = qw.rk4_heatmap(prob, t, b)
    po, no = c_oracle.heatmap(8192, "circular", "cerf_3", t, b, n_steps=96)
    assert np.abs(p - po).max() <= P_TOL and np.argmax(p) == np.argmax(po)
    lo, _ = qw.rk4_heatmap(prob, t, b, rows=(2, 5))
    assert np.array_equal(lo, p[2:5])
    # the streaming kernel and the half-ring reduction on the same grid
    for flags in (2, 8192):
        q, _ = qw.rk4_heatmap(qw.Problem(n=8192, topology="circular", schedule="cerf_3",
                                         n_steps=96, flags=flags), t, b)
        assert np.abs(q - p).max() <= P_TOL


def test_cluster_long_run_keeps_the_halo_protocol_in_step(qw):
    """Thousands of levels: the two-parity mailbox across CTAs must never slip a phase."""
    n, S = 8192, 3000
    T, g = np.array([700.0, 350.0]), np.array([1.2, 0.4])
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S)
    p, norm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(n, "circular", "lin", T, g, n_steps=S)
    assert np.abs(p - po).max() <= P_TOL and np.abs(norm - no).max() <= P_TOL
