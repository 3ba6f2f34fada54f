"""GPU: the drop-in modules, called the way the reference's scripts call their own
functions (module globals, [beta, T] arguments, negative probabilities, file names)."""

import importlib

import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu
PKG = "quantum-walks-time-dependent-hamiltonians_b200"


def _mod(name):
    return importlib.import_module(PKG + "." + name)


def _site(n):
    w = np.zeros((n, 1))
    w[int((n - 1) / 2)][0] = 1
    return w


def test_bcb_evaluate_probability_against_golden(golden):
    meta, arr = golden
    bcb = _mod("BCB")
    for c in meta["rk4_L0"]:
        bcb.dimension, bcb.topology, bcb.step_function = c["n"], c["topology"], c["schedule"]
        bcb.n_steps, bcb.step_size = c["rule"].get("n_steps"), c["rule"].get("step_size", 1e-3)
        val = bcb.evaluate_probability([c["beta"], c["T"]], _site(c["n"]))
        assert val.shape == (1,) and val[0] <= 0
        assert abs(-val[0] - c["p"]) <= 1e-10, c
        psi = bcb.solve_schrodinger_equation(c["T"], c["beta"])
        ref = arr["L0_psi_%d" % c["key"]]
        assert np.abs(psi - ref / np.sqrt(c["norm"])).max() <= 1e-11      # BCB normalises
        assert np.vdot(psi, psi).real == pytest.approx(1.0, abs=1e-13)
    bcb.n_steps, bcb.step_size = None, 1e-3


def test_default_step_rule_tracks_the_shipped_rk45(golden):
    """With the stub's default step_size = 1e-3 the drop-in reproduces what the reference
    prints today (RK45, tol 1e-6) to the accuracy of RK45 itself."""
    meta, _ = golden
    bcb = _mod("BCB")
    bcb.n_steps, bcb.step_size = None, 1e-3
    for c in meta["rk45"]:
        bcb.dimension, bcb.topology, bcb.step_function = c["n"], c["topology"], c["schedule"]
        val = bcb.evaluate_probability([c["beta"], c["T"]], _site(c["n"]))
        tol = 2e-4 if c["schedule"] in ("sqrt", "cbrt") else 3e-5
        assert abs(-val[0] - c["p"]) <= tol, c


def test_schroedinger_solver_module(golden):
    meta, arr = golden
    ss, rke = _mod("Schroedinger_Solver"), _mod("Runge_Kutta_Error")
    hm = [c for c in meta["heatmaps"] if c["kind"] == "RK45"][0]
    ss.dimension, ss.step_function, ss.n_steps, ss.step_size = 5, 1, None, 1e-3
    p = ss.grid_probability_evaluation(5, 1, 20, 4, 0.1, 5, 3, 0.92, 7)
    assert p.shape == (3, 4) and np.abs(p - arr[hm["key"]]).max() <= 3e-5
    ss.dimension, ss.step_function, ss.n_steps = 17, 3, 1024           # 3 = cbrt (SS:59)
    psi = ss.solve_schrodinger_equation(12.0, 2.0)
    c = [c for c in meta["rk4_L0"] if c["n"] == 17][0]
    assert np.abs(psi - arr["L0_psi_%d" % c["key"]]).max() <= 1e-11       # SS returns raw psi
    val = ss.evaluate_probability([2.0, 12.0], _site(17))
    assert abs(-val[0] - c["p"]) <= 1e-10
    other = np.zeros((17, 1))
    other[3][0] = 1                                                     # another projector
    val = ss.evaluate_probability([2.0, 12.0], other)
    assert val.shape == (1,)
    assert abs(-val[0] - abs(psi[3]) ** 2 / c["norm"]) <= 1e-12
    rke.dimension, rke.step_function, rke.n_steps = 17, 3, 1024
    assert abs(rke.solve_schrodinger_equation(12.0, 2.0) - c["norm"]) <= 1e-10
    assert abs(rke.norm_defect(12.0, 2.0) - (1 - c["norm"])) <= 1e-10
    assert ss.single_evaluation_benchmark([2.0, 12.0]) == 1
    ss.n_steps = rke.n_steps = None


def test_rhs_and_time_independent_comparator(golden):
    meta, arr = golden
    bcb = _mod("BCB")
    for c in meta["rhs"]:
        bcb.dimension, bcb.topology, bcb.step_function = c["n"], c["topology"], c["schedule"]
        dy = bcb.schrodinger_equation(c["t"], arr["rhs_y_%d" % c["key"]], c["beta"], c["T"])
        ref = arr["rhs_dy_%d" % c["key"]]
        assert isinstance(dy, list) and len(dy) == c["n"]
        assert np.abs(np.array(dy) - ref).max() <= 1e-12 * np.abs(ref).max()
    bcb.topology = "circular"
    for c in meta["static"]:
        bcb.dimension, bcb.n_steps = c["n"], c["n_steps"]
        bcb.ti_method = "exact"                      # Chebyshev propagator vs the reference expm
        val = bcb.evaluate_time_independent_probability(c["time"], c["gamma"], _site(c["n"]))
        assert val.shape == (1, 1)
        assert abs(-val[0, 0] - c["p_expm"]) <= 1e-12
        bcb.ti_method = "rk4"                        # fixed-step RK4 on the static Hamiltonian
        val = bcb.evaluate_time_independent_probability(c["time"], c["gamma"], _site(c["n"]))
        assert abs(-val[0, 0] - c["p_expm"]) <= 1e-8
        assert abs(-val[0, 0] - c["p_rk4"] * c["norm_rk4"]) <= 1e-10
    bcb.n_steps, bcb.ti_method = None, "exact"


def test_grid_eval_and_type_of_qw():
    bcb = _mod("BCB")
    bcb.dimension, bcb.topology, bcb.step_function, bcb.n_steps = 21, "circular", "cerf_3", 128
    t, b = np.linspace(0.1, 21, 6), np.linspace(0, 2.5, 4)
    bcb.type_of_qw = "dependent"
    p = bcb.grid_eval(t, b)
    po, _ = c_oracle.heatmap(21, "circular", "cerf_3", t, b, n_steps=128)
    assert p.shape == (4, 6) and np.abs(p - po).max() <= 1e-10
    bcb.type_of_qw, bcb.ti_method = "independent", "rk4"
    p = bcb.grid_eval(t, b)
    po, no = c_oracle.heatmap(21, "circular", "static", t, b, n_steps=128)
    assert np.abs(p - po).max() <= 1e-10
    bcb.ti_method = "exact"
    p = bcb.grid_eval(t, b)
    from oracle import rk4_numpy as rk
    pe = np.array([[rk.ti_exact(21, tt, bb)[0] for tt in t] for bb in b])
    assert p.shape == (4, 6) and np.abs(p - pe).max() <= 1e-12
    bcb.type_of_qw = "dependent"
    bcb.n_steps = None


def test_parallel_routine_file_contract_and_robustness(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    bcb, bl, rob = _mod("BCB"), _mod("BCB_latest"), _mod("Robustness")
    bcb.dimension, bcb.topology, bcb.step_function, bcb.type_of_qw = 9, "circular", "lin", "dependent"
    bcb.n_steps = 64
    assert bcb.parallel_routine(0.1, 9, 0, 2) == 0
    p = np.load("9_circular_probability_lin.npy")
    t, b = np.load("9_circ_time.npy"), np.load("9_circ_beta.npy")
    assert p.shape == (24, 24) and t.shape == (24,) and b.shape == (24,)
    po, _ = c_oracle.heatmap(9, "circular", "lin", t, b, n_steps=64)
    assert np.abs(p - po).max() <= 1e-10
    bcb.n_steps = None

    bl.dimension, bl.step_function, bl.type_of_qw, bl.n_steps = 11, "sqrt", "dependent", 64
    assert bl.parallel_routine() == 0
    p = np.load("11_probability_sqrt.npy")
    assert p.shape == (75, 4)
    t, b = np.load("11_circ_time.npy"), np.load("11_circ_beta.npy")
    assert t[2] == (np.pi / 2) * np.sqrt(11) and b.tolist() == bl.PRODUCTION_BETA
    po, _ = c_oracle.heatmap(11, "circular", "sqrt", t, b, n_steps=64)
    assert np.abs(p - po).max() <= 1e-10
    dim, r2, r4 = rob.routine(11, 2)                       # flag 2 = sqrt (Robustness.py:13)
    j = int(np.argmax(po[:, 2]))
    assert r2 == round(float(((po[j, 2] - po[j + 2, 2]) + (po[j, 2] - po[j - 2, 2])) / 2), 4)
    bl.n_steps = None


def test_qw_search_state_evolve():
    qws = _mod("QW_Search")
    qws.n_steps, qws.step_size = 256, None
    HC = qws.Hamiltonian(16, 1, 1, 7, 1, 1.0)              # time-dependent, cycle, lin
    HC.build_laplacian()
    HC.build_hamiltonian()
    psi = qws.State(16)
    psi.set_initial(0)
    psi.set_hamiltonian(HC)
    psi.evolve(16.0)
    po, no, psio, _ = c_oracle.rk4_batch(16, "circular", "lin", [16.0], [1.0], n_steps=256,
                                         return_psi=True)
    assert np.abs(psi.ket - psio[0]).max() <= 1e-11
    assert abs(psi.success_probability() - po[0]) <= 1e-10
    HI = qws.Hamiltonian(5, 0, 0, 1, 0, 5.0)               # time-independent, complete
    HI.build_laplacian()
    HI.build_hamiltonian()
    st = qws.State(5)
    st.set_initial(0)
    st.set_hamiltonian(HI)
    st.evolve(np.pi / (2 * np.sqrt(5)))
    assert abs(st.success_probability() - 1.0) <= 1e-9     # thesis Eq. 1.28
    qws.n_steps, qws.step_size = None, 1e-3


def test_robustness_ensemble_stats():
    rob = _mod("Robustness")
    qw = importlib.import_module(PKG)
    rng = np.random.default_rng(20200912)
    T0, g0 = (np.pi / 2) * 16.0, 1.0
    T = np.clip(T0 * (1 + 0.05 * rng.standard_normal(512)), 0, None)
    g = np.clip(g0 * (1 + 0.05 * rng.standard_normal(512)), 0, None)
    prob = qw.Problem(n=256, topology="circular", schedule="lin", n_steps=128)
    st = rob.ensemble(prob, T, g, p_reference=0.5)
    po, _, _ = c_oracle.rk4_batch(256, "circular", "lin", T, g, n_steps=128)
    assert st["mean"] == pytest.approx(po.mean(), abs=1e-11)
    assert st["std"] == pytest.approx(po.std(), abs=1e-11)
    assert st["R_sampled"] == pytest.approx(0.5 - po.mean(), abs=1e-11)


def test_integration_md_stub_runs():
    """The ctypes stub printed in INTEGRATION.md, executed verbatim against the built .so."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = [b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S)
             if "qw_b200_stub" in b][0]
    so = os.path.join(root, PKG, "libqw_b200.so")
    ns = {}
    exec(block.replace('"libqw_b200.so"', repr(so)), ns)
    t, b = np.linspace(0.1, 16, 5), np.linspace(0, 2.5, 4)
    p = ns["grid_eval"](16, "circular", "cerf_3", t, b, step_size=1e-2)
    po, _ = c_oracle.heatmap(16, "circular", "cerf_3", t, b, step_size=1e-2)
    assert p.shape == (4, 5) and np.abs(p - po).max() <= 1e-10
    with pytest.raises(ValueError):
        ns["grid_eval"](2, "circular", "lin", t, b)


def test_batched_optimisation_and_scipy_drop_in():
    """Row f3: the optimisers of Schroedinger_Solver.py:283-361.  The batched search must
    find at least the best cell of a dense oracle scan; the SciPy-driven drop-in (same
    calls as the reference, GPU objective) must agree with it."""
    ss = _mod("Schroedinger_Solver")
    ss.dimension, ss.step_function, ss.n_steps, ss.step_size = 5, 1, 200, None
    bnds = ([0.1, 5.0], [1.0, 20.0])
    res, loc, adi = ss.optimization(bnds, 'GPU', 0)
    t, b = np.linspace(1, 20, 60), np.linspace(0.1, 5, 60)
    po, _ = c_oracle.heatmap(5, "circular", "lin", t, b, n_steps=200)
    assert res[0] == 5 and res[1] >= po.max() - 1e-9 and res[1] <= 1.0
    assert 0.1 <= res[2] <= 5.0 and 1.0 <= res[3] <= 20.0
    pchk, _, _ = c_oracle.rk4_batch(5, "circular", "lin", [res[3]], [res[2]], n_steps=200)
    assert abs(pchk[0] - res[1]) <= 1e-10                       # the reported optimum is real
    assert loc.shape[1] == 3 and (np.diff(loc[:, 0]) <= 0).all()
    res2, loc2, _ = ss.optimization(bnds, 'SHGO', 0)
    assert res2[1] <= res[1] + 1e-6 and res2[1] > 0.9           # shgo finds a good local maximum
    res3 = ss.optimization(bnds, 'DUAL', 30)
    assert abs(res3[1] - res[1]) < 5e-3
    assert ss.optimization_precision(None, -0.995, None) and not ss.optimization_precision(
        None, -0.5, None)
    ss.n_steps, ss.step_size = None, 1e-3


def test_integrator_rk45_reproduces_reference_printouts(golden):
    """With ``integrator = 'rk45'`` the drop-in prints what the reference prints today."""
    meta, arr = golden
    bcb, ss = _mod("BCB"), _mod("Schroedinger_Solver")
    bcb.integrator, bcb.rtolerance, bcb.atolerance = "rk45", 1e-6, 1e-6
    for c in meta["rk45"]:
        bcb.dimension, bcb.topology, bcb.step_function = c["n"], c["topology"], c["schedule"]
        val = bcb.evaluate_probability([c["beta"], c["T"]], _site(c["n"]))
        assert val.shape == (1,) and abs(-val[0] - c["p"]) <= 1e-9, c
    bcb.integrator, bcb.topology = "rk4", "circular"
    hm = [c for c in meta["heatmaps"] if c["kind"] == "RK45"][0]
    ss.dimension, ss.step_function, ss.integrator = 5, 1, "rk45"
    p = ss.grid_probability_evaluation(5, 1, 20, 4, 0.1, 5, 3, 0.92, 7)
    assert np.abs(p - arr[hm["key"]]).max() <= 1e-9
    ss.integrator = "rk4"


def test_runge_kutta_error_main_block():
    """The __main__ of Runge_Kutta_Error.py (RKE:463-532): norm defect 1 - <psi|psi> of RK45
    at N = 51, beta = 0.5, T = 300, for atol in {1e-6, 1e-7, 1e-8} (rtol = 1e-6)."""
    from oracle import rk4_numpy as rk
    rke = _mod("Runge_Kutta_Error")
    rke.dimension, rke.step_function, rke.integrator = 51, 1, "rk45"
    for atol in (1e-6, 1e-7, 1e-8):
        rke.rtolerance, rke.atolerance = 1e-6, atol
        defect = rke.norm_defect(300.0, 0.5)
        _, norm, _, _ = rk.rk45_scipy(51, "circular", "lin", 300.0, 0.5, rtol=1e-6, atol=atol)
        assert abs(defect - (1 - norm)) <= 1e-9, (atol, defect, 1 - norm)
        assert abs(defect) < 1e-3                    # "error in the order of <10^-4" (SS:112)
    rke.integrator, rke.rtolerance, rke.atolerance = "rk4", 1e-6, 1e-6


def test_production_sweep_on_streams_writes_the_same_files(tmp_path, monkeypatch):
    """BCB_latest.production_sweep: the fused multi-problem launch and the multi-stream issue
    order must give the files of the reference's sequential loop (names of
    BCB_latest.py:287-293), bit for bit -- with the half-ring reduction (default) and without."""
    bl = _mod("BCB_latest")
    for mirror in (True, False):
        out = {}
        for mode, kw in (("seq", dict(streams=1, fused=False)), ("streams", dict(streams=8, fused=False)),
                         ("fused", dict())):
            d = tmp_path / ("%s_%d" % (mode, mirror))
            d.mkdir()
            monkeypatch.chdir(d)
            bl.integrator, bl.n_steps, bl.step_size, bl.type_of_qw = "rk4", None, 1e-2, "dependent"
            bl.topology, bl.mirror = "circular", mirror
            bl.production_sweep(step_functions=("lin", "cerf_3"), dims=(5, 21, 23, 51, 64), **kw)
            out[mode] = {f.name: np.load(f) for f in sorted(d.glob("*.npy"))}
        assert sorted(out["seq"]) == sorted(out["fused"]) == sorted(out["streams"])
        assert len(out["seq"]) == 2 * 5 + 2 * 5
        for name in out["seq"]:
            assert np.array_equal(out["seq"][name], out["fused"][name]), (mirror, name)
            assert np.array_equal(out["seq"][name], out["streams"][name]), (mirror, name)
        assert out["fused"]["51_probability_lin.npy"].shape == (75, 4)
    bl.step_size, bl.mirror = 1e-3, True


def test_mirror_knob_of_the_drop_in_modules():
    """BCB.mirror = True routes grid_eval through the mirror-reduced kernel: same heatmap."""
    bcb = _mod("BCB")
    bcb.dimension, bcb.step_function, bcb.topology = 256, "lin", "circular"
    bcb.type_of_qw, bcb.n_steps, bcb.step_size = "dependent", 200, None
    t, b = np.linspace(0.1, 30.0, 5), np.linspace(0.0, 2.0, 6)
    bcb.mirror = False
    full = bcb.grid_eval(t, b)
    bcb.mirror = True
    half = bcb.grid_eval(t, b)
    assert np.abs(full - half).max() <= 1e-13 and np.argmax(full) == np.argmax(half)
    # complete graph (BCB.py:30-37): the (psi_w, psi_other) pair instead of 512 amplitudes
    bcb.dimension, bcb.topology, bcb.n_steps = 512, "complete", 600
    t, b = np.linspace(0.01, 0.3, 5), np.linspace(0.0, 900.0, 6)
    bcb.mirror = False
    full = bcb.grid_eval(t, b)
    bcb.mirror = True
    pair = bcb.grid_eval(t, b)
    bcb.mirror, bcb.n_steps, bcb.step_size, bcb.topology = True, None, 1e-3, "circular"
    assert np.isfinite(full).all() and full.max() > 0.1
    assert np.abs(full - pair).max() <= 1e-12 and np.argmax(full) == np.argmax(pair)
