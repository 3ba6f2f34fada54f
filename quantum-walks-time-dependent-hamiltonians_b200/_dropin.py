"""Shared machinery of the drop-in modules (Schroedinger_Solver, Runge_Kutta_Error, BCB,
BCB_latest).

The reference configures itself through module-level globals read at call time
(``dimension``, ``step_function``, ``topology``, ``type_of_qw``), so
each drop-in module calls ``install(globals(), flavor)`` and gets functions with the
reference's names and argument meaning that read *that module's* globals on every call.
New knobs live beside the old ones: ``n_steps`` / ``step_size`` (fixed-step rule,
QW_Search.py:149-164; default step_size = 1e-3 as in the stub), ``gamma_on``, ``device``.
``integrator`` selects the fixed-step RK4 of the north star ('rk4', default; ``rtolerance`` /
``atolerance`` unused) or SciPy's adaptive RK45 as the reference ships it ('rk45', tolerances
honoured -- BASELINE.md section 1).

Everything that integrates goes to the CUDA kernels through ``engine``; nothing here
falls back to the CPU.
"""

import time as _time

import numpy as np

from . import engine

_SS_SCHEDULES = {1: "lin", 2: "sqrt", 3: "cbrt"}      # Schroedinger_Solver.py:55-60


def _schedule_name(g, flavor):
    s = g["step_function"]
    if flavor in ("SS", "RKE"):
        if s in _SS_SCHEDULES:
            return _SS_SCHEDULES[s]
        if isinstance(s, str):
            return s
        raise ValueError("step_function value not defined: %r" % (s,))
    return s


def _topology(g, flavor):
    return "circular" if flavor in ("SS", "RKE") else g.get("topology", "circular")


FIXED_GEOMETRY = 32768        # QW_FLAG_FIXED_GEOMETRY


def _problem(g, flavor, schedule=None, topology=None, dimension=None, extra_flags=0):
    n_steps, step_size = g.get("n_steps"), g.get("step_size")
    if n_steps is not None:
        step_size = None
    dim = g.get("dimension") if dimension is None else dimension
    if dim is None:
        raise ValueError("module global 'dimension' is not set")
    return engine.Problem(
        n=int(dim), topology=_topology(g, flavor) if topology is None else topology,
        schedule=_schedule_name(g, flavor) if schedule is None else schedule,
        gamma_on=g.get("gamma_on", "oracle"), n_steps=n_steps, step_size=step_size,
        flags=(8192 if g.get("mirror", True) else 0) | extra_flags)   # QW_FLAG_MIRROR: exact
                                                           # symmetry reduction for the uniform start


def _marked_vertex(n):
    return int((n - 1) / 2)                               # BCB.py:96-97


def laplacian(dimension, topology):
    """Dense graph Laplacian (inspection helper; the integrator never builds it)."""
    n = int(dimension)
    if topology == "complete":
        return n * np.eye(n) - np.ones((n, n))
    if topology in ("circular", "cycle"):
        if n < 3:
            raise ValueError("cycle graph needs dimension >= 3")
        lap = 2.0 * np.eye(n)
        i = np.arange(n)
        lap[i, (i + 1) % n] = -1.0
        lap[i, (i - 1) % n] = -1.0
        return lap
    raise ValueError("undefined topology: %r" % (topology,))


def schedule_value(tau, name):
    """g_T(t/T) on the host, for the dense-matrix helpers only."""
    if name == "lin":
        return tau
    if name == "sqrt":
        return np.sqrt(tau)
    if name == "cbrt":
        return np.cbrt(tau)
    if name in ("cerf_3", "cerf_5", "cerf_7"):
        return 0.5 * (1 + (2 * tau - 1) ** int(name[-1]))
    raise ValueError("step_function value not defined: %r" % (name,))


def install(g, flavor):
    """Define the reference's functions in module namespace ``g``."""

    def generate_hamiltonian(dimension, beta, time, T):
        """Dense H(t) with the reference's conventions (BCB.py:28-99, SS:21-70).  Kept for
        inspection and API compatibility; the kernels apply H implicitly."""
        lap = laplacian(dimension, _topology(g, flavor))
        zero_guard = (time == 0) if flavor in ("SS", "RKE") else (time == 0 or T == 0)
        if zero_guard:
            return lap
        s = schedule_value(float(time) / T, _schedule_name(g, flavor))
        w = _marked_vertex(dimension)
        if g.get("gamma_on", "oracle") == "laplacian":
            ham = (1 - s) * beta * lap
            ham[w, w] += -s
        else:
            ham = (1 - s) * lap
            ham[w, w] += -s * beta
        return ham

    def schrodinger_equation(t, y, beta, T):
        """d psi/dt = -i H(t) psi as a list of N complex (BCB.py:163-174) -- one launch
        of the implicit-operator kernel."""
        dy = engine.rhs(_problem(g, flavor), t, T, beta, np.asarray(y, dtype=complex),
                        device=g.get("device"))
        return list(dy)

    def _rk45():
        """integrator = 'rk45': the reference's shipped SciPy RK45 (rtolerance / atolerance
        honoured) instead of the fixed-step RK4 of the north star."""
        return g.get("integrator", "rk4") == "rk45"

    def _integrate(T, beta, want_psi, schedule=None, topology=None):
        prob = _problem(g, flavor, schedule, topology)
        if _rk45():
            return engine.rk45_batch(prob, [float(T)], [float(beta)],
                                     rtol=g.get("rtolerance", 1e-6), atol=g.get("atolerance", 1e-6),
                                     return_psi=want_psi, device=g.get("device"))
        return engine.rk4_batch(prob, [float(T)], [float(beta)], return_psi=want_psi,
                                device=g.get("device"))

    def solve_schrodinger_equation(time, beta):
        """psi(T): SS returns it raw (SS:102), BCB normalised (BCB:189-190),
        Runge_Kutta_Error returns <psi|psi>.real (RKE:103)."""
        p, norm, psi = _integrate(time, beta, True)
        if flavor == "RKE":
            return float(norm[0])
        if flavor in ("BCB", "BCBL"):
            return psi[0] * (1 / np.sqrt(norm[0]))
        return psi[0]

    def evaluate_probability(x, oracle_site_state):
        """-|<w|psi(T)>|^2 / <psi|psi> for x = [beta, T], shape (1,) (BCB.py:193-203)."""
        n = int(g["dimension"])
        osite = np.asarray(oracle_site_state, dtype=float).reshape(n, -1)
        basis = np.zeros((n, 1))
        basis[_marked_vertex(n)][0] = 1
        if osite.shape == (n, 1) and np.array_equal(osite, basis):
            p, norm = _integrate(x[1], x[0], False)          # fused epilogue
            prob = np.array([p[0]])
        else:                                                # projection on another state
            p, norm, psi = _integrate(x[1], x[0], True)
            prob = np.abs(np.dot(osite.transpose(), psi[0] / np.sqrt(norm[0]))) ** 2
        if (prob > 1).any():
            print('Error: probability out of bounds: ', prob)
            return None
        return -prob

    def generate_time_independent_hamiltonian(dimension, gamma):
        """L_cycle - gamma |w><w| (BCB.py:101-135; always the cycle graph)."""
        ham = laplacian(dimension, "circular")
        w = _marked_vertex(dimension)
        ham[w][w] += -gamma
        return ham

    def evaluate_time_independent_probability(time, gamma, oracle_site_state):
        """-|<w|exp(-iHt)|s>|^2, shape (1, 1) (BCB.py:137-159).  Argument order is
        (time, gamma); always the cycle graph, as in the reference.  ``ti_method='exact'``
        (default) evaluates the propagator by a Chebyshev expansion on the device, matching
        the reference's scipy expm; ``'rk4'`` integrates schedule 'static' with the fixed-step
        kernel instead."""
        if g.get("ti_method", "exact") == "exact":
            prob = _problem(g, flavor, schedule="static", topology="circular")
            p, norm = engine.ti_exact_batch(prob, [float(time)], [float(gamma)],
                                            device=g.get("device"))
            return -np.array([[p[0]]])
        p, norm = _integrate(time, gamma, False, schedule="static", topology="circular")
        return -np.array([[p[0] * norm[0]]])          # the reference does not renormalise

    def heatmap2d(*args, **kwargs):
        """Plotting is outside the accelerated path: signature kept,
        nothing is drawn."""
        return None

    def _heatmap(time_array, beta_array, schedule=None, topology=None, dimension=None):
        prob = _problem(g, flavor, schedule, topology, dimension)
        t = np.asarray(time_array, dtype=np.float64)
        b = np.asarray(beta_array, dtype=np.float64)
        if _rk45():
            p, _ = engine.rk45_heatmap(prob, t, b, rtol=g.get("rtolerance", 1e-6),
                                       atol=g.get("atolerance", 1e-6), device=g.get("device"))
        else:
            p, _ = engine.rk4_heatmap(prob, t, b, device=g.get("device"))
        return p

    def grid_eval(time_array, beta_array):
        """probability[beta][T] (BCB.py:239-259).  'independent' integrates the static
        Hamiltonian (the reference's branch passes the wrong arguments, BCB.py:257)."""
        kind = g.get("type_of_qw", "dependent")
        if kind == "dependent":
            return _heatmap(time_array, beta_array)
        if kind == "independent":
            if g.get("ti_method", "exact") == "exact":
                p, _ = engine.ti_exact_heatmap(
                    _problem(g, flavor, schedule="static", topology="circular"),
                    np.asarray(time_array, dtype=np.float64),
                    np.asarray(beta_array, dtype=np.float64), device=g.get("device"))
                return p
            return _heatmap(time_array, beta_array, schedule="static", topology="circular")
        raise ValueError("undefined type_of_qw: %r" % (kind,))

    def grid_probability_evaluation(dimension, time_lb, time_up, time_sampling_points,
                                    beta_lb, beta_up, beta_sampling_points, non_time=None,
                                    non_prob=None):
        """SS:426-444 -- one kernel launch instead of the double loop."""
        time_array = np.linspace(time_lb, time_up, time_sampling_points)
        beta_array = np.linspace(beta_lb, beta_up, beta_sampling_points)
        p = _heatmap(time_array, beta_array, dimension=dimension)
        g["heatmap2d"](p, time_array, beta_array, non_time, non_prob)
        return p

    def single_evaluation_benchmark(x):
        """SS:446-457."""
        n = int(g["dimension"])
        osite = np.zeros((n, 1))
        osite[_marked_vertex(n)][0] = 1
        tic = _time.perf_counter()
        g["evaluate_probability"](x, osite)
        toc = _time.perf_counter() - tic
        print(n, ": ", int(toc))
        print()
        return 1

    # ---- adiabatic-theorem diagnostics (Schroedinger_Solver.py:146-280) -------------------
    def _slope(s, name, reference):
        """g'(s); the reference's cube-root slope is 1/(3 cbrt(s)) (SS:204-206)."""
        if name == "lin":
            return 1.0
        if name == "sqrt":
            return 1.0 / (2.0 * np.sqrt(s))
        if name == "cbrt":
            return 1.0 / (3.0 * np.cbrt(s)) if reference else 1.0 / (3.0 * np.cbrt(s) ** 2)
        if name in ("cerf_3", "cerf_5", "cerf_7"):
            k = int(name[-1])
            return k * (2.0 * s - 1.0) ** (k - 1)
        raise ValueError("step_function value not defined: %r" % (name,))

    def generate_hamiltonian_derivative(s, beta, derivative):
        """Dense H(s) (derivative = 0) or dH/ds (derivative = 1) with the reference's
        conventions (SS:146-207).  Inspection helper; the diagnostics below never build it."""
        n = int(g["dimension"])
        lap = laplacian(n, _topology(g, flavor))
        name = _schedule_name(g, flavor)
        w = _marked_vertex(n)
        if derivative == 0:
            gs = schedule_value(s, name)
            ham = (1 - gs) * lap
            ham[w, w] += -gs * beta
        elif derivative == 1:
            sl = _slope(s, name, g.get("reference_derivative", True))
            ham = -sl * lap
            ham[w, w] += -sl * beta
        else:
            raise ValueError("derivative flag unknown: %r" % (derivative,))
        return ham

    def _adiabatic(s, beta):
        gap, cpl, e0 = engine.adiabatic_batch(
            _problem(g, flavor), np.atleast_1d(float(s)), np.atleast_1d(float(beta)),
            reference_derivative=g.get("reference_derivative", True), device=g.get("device"))
        return float(gap[0]), float(cpl[0])

    def compute_energy_diff(s, beta):
        """E1 - E0 of H(s) (SS:248-253) from the secular equation on the device."""
        return _adiabatic(s, beta)[0]

    def compute_gamma(s, beta):
        """-|<phi1| dH/ds |phi0>| as a (1, 1) array (SS:225-246)."""
        return -np.abs(np.array([[_adiabatic(s, beta)[1]]]))

    def adiabatic_theorem_check(current_par):
        """[adiabatic_time, gamma_max, energy_min, flag, minutes] (SS:255-280): the two shgo
        runs over dense eigensolves become a few batched launches (engine.adiabatic_check)."""
        tic = _time.perf_counter()
        res = engine.adiabatic_check(_problem(g, flavor), float(current_par[2]),
                                     reference_derivative=g.get("reference_derivative", True),
                                     device=g.get("device"))
        flag = "True" if res["energy_min"] > 0 else "False"
        return [res["adiabatic_time"], res["gamma_max"], res["energy_min"], flag,
                (tic - _time.perf_counter()) / 60]

    def optimization_precision(x, probability, context):
        """Callback of SS:125-130: stop once p >= 0.99."""
        return bool(probability <= -0.99)

    def arrange_local_maxima(results):
        """SS:133-143."""
        local_maxima = np.empty([len(results.funl), 3])
        for i in range(len(results.funl)):
            local_maxima[i][0] = -results.funl[i]
            local_maxima[i][1] = results.xl[i][0]
            local_maxima[i][2] = results.xl[i][1]
        return local_maxima

    def optimization(par_bnds, optimization_method, its):
        """Maximise p over (beta, T) in ``par_bnds`` (SS:283-361).  'SHGO', 'BH' and 'DUAL'
        drive the same SciPy optimisers as the reference with the GPU objective (one point
        per call); 'GPU' evaluates whole candidate sets per launch
        (``engine.maximize_probability``).  Return values as in the reference."""
        n = int(g["dimension"])
        osite = np.zeros((n, 1))
        osite[_marked_vertex(n)][0] = 1
        tic = _time.perf_counter()
        if optimization_method == 'GPU':
            p, beta, T, loc = engine.maximize_probability(
                _problem(g, flavor), par_bnds[0], par_bnds[1], device=g.get("device"))
            comp_time = int(_time.perf_counter() - tic) / 60
            return [n, p, beta, T, comp_time], loc, [0, 0, 0, 0, 0]
        from scipy.optimize import basinhopping, dual_annealing, shgo
        objective = g["evaluate_probability"]
        if optimization_method == 'SHGO':
            res = shgo(objective, par_bnds, n=40, iters=1, args=(osite,),
                       sampling_method='sobol')
            comp_time = int(_time.perf_counter() - tic) / 60
            return ([n, float(-res.fun), res.x[0], res.x[1], comp_time],
                    arrange_local_maxima(res), [0, 0, 0, 0, 0])
        if optimization_method == 'BH':
            kwargs = dict(method="L-BFGS-B", bounds=par_bnds, args=osite)
            res = basinhopping(objective, np.array([1., 10]), minimizer_kwargs=kwargs,
                               niter=its)
            comp_time = int(_time.perf_counter() - tic) / 60
            return [n, float(-res.fun), res.x[0], res.x[1], comp_time]
        if optimization_method == 'DUAL':
            res = dual_annealing(objective, par_bnds, args=(osite,), maxiter=its)
            comp_time = int(_time.perf_counter() - tic) / 60
            return [n, float(-res.fun), res.x[0], res.x[1], comp_time]
        raise ValueError('invalid optimization method: %r' % (optimization_method,))

    def _save_and_report(prob, time_array, beta, file_probability, tic):
        import torch.distributed as dist
        n = g["dimension"]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0:
            np.save(file_probability, prob)
            np.save(str(n) + '_circ_time.npy', time_array)
            np.save(str(n) + '_circ_beta.npy', beta)
        toc = _time.perf_counter() - tic
        print('Success: N ', n, ' in ', int(toc / 60), ' min')
        return 0

    def _sharded(time_array, beta):
        """This rank's block of beta rows on its GPU (CUDA tensors all the way into the
        collective: NCCL moves device memory only), same integrator selection as grid_eval."""
        import torch
        kind = g.get("type_of_qw", "dependent")
        if kind == "dependent":
            # fixed geometry: the files do not depend on the number of GPUs the rows are cut over
            prob = _problem(g, flavor, extra_flags=FIXED_GEOMETRY)
            if _rk45():
                def fn(problem, t, b, rows, dev):
                    return engine.rk45_heatmap(problem, t, b, rtol=g.get("rtolerance", 1e-6),
                                               atol=g.get("atolerance", 1e-6), rows=rows,
                                               device=dev)
            else:
                def fn(problem, t, b, rows, dev):
                    return engine.rk4_heatmap(problem, t, b, rows=rows, device=dev)
        elif kind == "independent":
            prob = _problem(g, flavor, schedule="static", topology="circular",
                            extra_flags=FIXED_GEOMETRY)
            if g.get("ti_method", "exact") == "exact":
                def fn(problem, t, b, rows, dev):
                    return engine.ti_exact_heatmap(problem, t, b, rows=rows, device=dev)
            else:
                def fn(problem, t, b, rows, dev):
                    return engine.rk4_heatmap(problem, t, b, rows=rows, device=dev)
        else:
            raise ValueError("undefined type_of_qw: %r" % (kind,))

        def compute(problem, t, b, rows):
            dev = engine._device(g.get("device"))
            return fn(problem, torch.from_numpy(np.ascontiguousarray(t)).to(dev),
                      torch.from_numpy(np.ascontiguousarray(b)).to(dev), rows, dev)

        p, _ = engine.heatmap_sharded(prob, time_array, beta, compute=compute)
        return p

    def parallel_routine_bcb(lb_time, ub_time, lb_beta, ub_beta):
        """BCB.py:262-315: the Ray fan-out over beta blocks becomes row shards over the
        GPUs of the process group (one process per GPU) and one all_gather."""
        tic = _time.perf_counter()
        cpu_count, time_sampling_points, sampling_per_cpu_count = 4, 24, 6
        beta = np.linspace(lb_beta, ub_beta, cpu_count * sampling_per_cpu_count)
        time_array = np.linspace(lb_time, ub_time, time_sampling_points)
        prob = _sharded(time_array, beta)
        n, topo = g["dimension"], g.get("topology", "circular")
        if g.get("type_of_qw", "dependent") == "dependent":
            name = str(n) + '_' + topo + '_probability_' + g["step_function"] + '.npy'
        else:
            name = str(n) + '_' + topo + '_probability_static.npy'
        return _save_and_report(prob, time_array, beta, name, tic)

    def parallel_routine_bcbl():
        """BCB_latest.py:240-297: the thesis' production grid (75 beta x 4 T)."""
        tic = _time.perf_counter()
        beta = list(g["PRODUCTION_BETA"])
        n = g["dimension"]
        time_array = [(np.pi / 4) * np.sqrt(n), np.sqrt(n), (np.pi / 2) * np.sqrt(n),
                      2 * np.sqrt(n)]
        prob = _sharded(time_array, beta)
        if g.get("type_of_qw", "dependent") == "dependent":
            name = str(n) + '_probability_' + g["step_function"] + '.npy'
        else:
            name = str(n) + '_probability_static.npy'
        return _save_and_report(prob, time_array, beta, name, tic)

    g.update(generate_hamiltonian=generate_hamiltonian,
             schrodinger_equation=schrodinger_equation,
             solve_schrodinger_equation=solve_schrodinger_equation,
             evaluate_probability=evaluate_probability, heatmap2d=heatmap2d)
    if flavor in ("SS", "RKE"):
        g.update(grid_probability_evaluation=grid_probability_evaluation,
                 single_evaluation_benchmark=single_evaluation_benchmark,
                 optimization=optimization, optimization_precision=optimization_precision,
                 arrange_local_maxima=arrange_local_maxima,
                 generate_hamiltonian_derivative=generate_hamiltonian_derivative,
                 compute_energy_diff=compute_energy_diff, compute_gamma=compute_gamma,
                 adiabatic_theorem_check=adiabatic_theorem_check)
    else:
        g.update(generate_time_independent_hamiltonian=generate_time_independent_hamiltonian,
                 evaluate_time_independent_probability=evaluate_time_independent_probability,
                 grid_eval=grid_eval,
                 parallel_routine=parallel_routine_bcb if flavor == "BCB"
                 else parallel_routine_bcbl)
