"""Host side of the B200 RK4 heatmap path: buffers, streams and sharding around the C ABI.

PyTorch is used for device memory, streams and ``torch.distributed`` only; all arithmetic
happens in the kernels of ``libqw_b200.so``.  Host arrays (numpy / CPU tensors) in -> numpy
out, with the H2D / D2H copies done here; CUDA tensors in -> CUDA tensors out, stream
ordered and without synchronisation.

Reference call sites this replaces: the double loop of ``grid_probability_evaluation``
(Schroedinger_Solver.py:438-441) / ``grid_eval`` (BCB.py:250-253) and the Ray fan-out of
``parallel_routine`` (BCB.py:279-292).
"""

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib


def _schedule_id(s):
    """The reference's three schedule encodings (SS ints 1/2/3 are handled by the SS
    shim; here: BCB strings, QW_Search ints)."""
    if isinstance(s, str):
        if s == "independent":
            s = "static"
        if s not in _lib.SCHEDULES:
            raise ValueError("step_function value not defined: %r" % (s,))
        return _lib.SCHEDULES[s]
    return int(s)


def _topology_id(t):
    if isinstance(t, str):
        try:
            return {"complete": _lib.TOPO_COMPLETE, "circular": _lib.TOPO_CYCLE,
                    "cycle": _lib.TOPO_CYCLE}[t]
        except KeyError:
            raise ValueError("undefined topology: %r" % (t,))
    return int(t)


def _gamma_on_id(g):
    if isinstance(g, str):
        try:
            return {"oracle": _lib.GAMMA_ON_ORACLE, "laplacian": _lib.GAMMA_ON_LAPLACIAN}[g]
        except KeyError:
            raise ValueError("undefined gamma_on: %r" % (g,))
    return int(g)


@dataclass
class Problem:
    """What the reference keeps in module globals (BCB.py:20-24) plus the step rule."""
    n: int
    topology: object = "circular"
    schedule: object = "lin"
    gamma_on: object = "oracle"
    target: Optional[int] = None
    n_steps: Optional[int] = None        # h = T / n_steps
    step_size: Optional[float] = None    # h fixed, n = int(T/h)   (QW_Search.py:153-154)
    flags: int = 0

    def c_struct(self, per_point_steps=False):
        """``per_point_steps``: the call carries its own n_steps array (or needs none)."""
        if self.n_steps is not None and self.step_size is not None:
            raise ValueError("give exactly one of n_steps / step_size")
        if self.n_steps is None and self.step_size is None and not per_point_steps:
            raise ValueError("give exactly one of n_steps / step_size")
        return _lib.QwProblem(
            n=int(self.n), topology=_topology_id(self.topology),
            schedule=_schedule_id(self.schedule), gamma_on=_gamma_on_id(self.gamma_on),
            flags=int(self.flags), target=-1 if self.target is None else int(self.target),
            step_size=float(self.step_size) if self.step_size is not None else 0.0,
            n_steps=int(self.n_steps) if self.n_steps is not None else 0)


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("qw_b200 needs a CUDA device (there is no CPU fallback)")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _as_dev(x, dev, dtype=torch.float64):
    """-> (contiguous CUDA tensor, was_on_host)."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x.to(dtype).contiguous(), False
        return x.to(dtype).contiguous().to(dev, non_blocking=True), True
    arr = np.ascontiguousarray(np.atleast_1d(np.asarray(x)),
                               dtype=np.float64 if dtype == torch.float64 else np.int64)
    return torch.from_numpy(arr).to(dev, non_blocking=True), True


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def longest_first_order(n_steps):
    """Launch order for unequal work per point (step_size rule: work ~ T)."""
    n_steps = np.asarray(n_steps)
    return np.argsort(-n_steps, kind="stable").astype(np.int64)


def rk4_batch(problem, T, gamma, n_steps=None, order=None, return_psi=False, device=None):
    """Integrate independent (T_q, gamma_q) points.  Returns (p, norm[, psi]).

    Host inputs (lists, numpy arrays) go straight to the host-buffer entry point of the library
    (``qw_rk4_batch_host``: persistent pinned staging, no framework tensors, and for a handful of
    points on a tiny graph the parallel-in-time kernels): the single-point calls of the reference
    (evaluate_probability, the SciPy optimisers) are latency bound."""
    dev = _device(device)
    lib = _lib.load()
    host_in = not any(isinstance(x, torch.Tensor) for x in (T, gamma, n_steps, order))
    if host_in and order is None:
        Th = np.ascontiguousarray(np.atleast_1d(np.asarray(T, dtype=np.float64)))
        gh = np.atleast_1d(np.asarray(gamma, dtype=np.float64))
        if gh.size == 1 and Th.size > 1:
            gh = np.full(Th.shape, gh[0])
        if Th.shape != gh.shape or Th.ndim != 1:
            raise ValueError("T and gamma must have the same length")
        if n_steps is not None and np.size(n_steps) != Th.size:
            raise ValueError("n_steps must have one entry per point")
        return rk4_batch_host(problem, Th, gh, n_steps=n_steps, return_psi=return_psi,
                              device=dev.index if dev.index is not None else 0)
    with torch.cuda.device(dev):
        dT, host = _as_dev(T, dev)
        dG, _ = _as_dev(gamma, dev)
        if dG.numel() == 1 and dT.numel() > 1:
            dG = dG.expand(dT.numel()).contiguous()
        if dT.numel() != dG.numel():
            raise ValueError("T and gamma must have the same length")
        P = dT.numel()
        dS = _as_dev(n_steps, dev, torch.int64)[0] if n_steps is not None else None
        if dS is not None and dS.numel() != P:
            raise ValueError("n_steps must have one entry per point")
        dO = _as_dev(order, dev, torch.int64)[0] if order is not None else None
        prob = problem.c_struct(per_point_steps=n_steps is not None)
        p = torch.empty(P, dtype=torch.float64, device=dev)
        norm = torch.empty(P, dtype=torch.float64, device=dev)
        psi = (torch.empty((P, int(problem.n)), dtype=torch.complex128, device=dev)
               if return_psi else None)
        _lib.check(lib.qw_rk4_batch_dev(ctypes.byref(prob), _ptr(dT), _ptr(dG), _ptr(dS),
                                        _ptr(dO), P, _ptr(p), _ptr(norm), _ptr(psi),
                                        _stream()))
        out = (p, norm) + ((psi,) if return_psi else ())
        if host:
            out = tuple(t.cpu().numpy() for t in out)
        return out


def rk4_heatmap(problem, time_array, beta_array, rows=None, device=None):
    """probability[beta][T] (and ||psi||^2) for rows [rows[0], rows[1]) of the beta axis,
    in the reference's layout (BCB.py:247, SS:436)."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        dT, host = _as_dev(time_array, dev)
        dB, _ = _as_dev(beta_array, dev)
        nT, nB = dT.numel(), dB.numel()
        b0, b1 = (0, nB) if rows is None else (int(rows[0]), int(rows[1]))
        p = torch.empty((b1 - b0, nT), dtype=torch.float64, device=dev)
        norm = torch.empty((b1 - b0, nT), dtype=torch.float64, device=dev)
        prob = problem.c_struct()
        _lib.check(lib.qw_rk4_grid_dev(ctypes.byref(prob), _ptr(dT), nT, _ptr(dB), nB, b0, b1,
                                       _ptr(p), _ptr(norm), _stream()))
        if host:
            return p.cpu().numpy(), norm.cpu().numpy()
        return p, norm


def rk4_heatmaps(problems, time_arrays, beta_arrays, device=None):
    """Several heatmaps in as few launches as possible (``qw_rk4_multi_grid_dev``): problems
    whose warp-packed kernel coincides share ONE grid, the rest are launched one by one on the
    same stream.  The reference's production sweep (BCB_latest.py:309-320) and the three
    schedules of BASELINE config 2 are such collections of small problems.  Returns a list of
    (p, norm) pairs, bitwise those of separate ``rk4_heatmap`` calls; host arrays in -> numpy
    out (one synchronisation for all of them), CUDA tensors in -> CUDA tensors out."""
    dev = _device(device)
    lib = _lib.load()
    n = len(problems)
    if not (len(time_arrays) == n and len(beta_arrays) == n):
        raise ValueError("one time axis and one beta axis per problem")
    with torch.cuda.device(dev):
        jobs = (_lib.QwGridJob * n)()
        keep, outs, host = [], [], False
        for k in range(n):
            dT, h1 = _as_dev(time_arrays[k], dev)
            dB, _ = _as_dev(beta_arrays[k], dev)
            host = host or h1
            nT, nB = dT.numel(), dB.numel()
            p = torch.empty((nB, nT), dtype=torch.float64, device=dev)
            norm = torch.empty((nB, nT), dtype=torch.float64, device=dev)
            keep += [dT, dB]
            outs.append((p, norm))
            j = jobs[k]
            j.prob = problems[k].c_struct()
            j.time_array, j.n_time = dT.data_ptr(), nT
            j.beta_array, j.n_beta, j.beta_begin, j.beta_end = dB.data_ptr(), nB, 0, nB
            j.p, j.norm = p.data_ptr(), norm.data_ptr()
        _lib.check(lib.qw_rk4_multi_grid_dev(jobs, n, _stream()))
        if host:
            torch.cuda.current_stream().synchronize()
            return [(p.cpu().numpy(), nrm.cpu().numpy()) for p, nrm in outs]
        return outs


def rk4_heatmap_host(problem, time_array, beta_array, rows=None, device=0):
    """Same through the host-pointer C entry point (copies inside the library)."""
    lib = _lib.load()
    t = np.ascontiguousarray(time_array, dtype=np.float64)
    b = np.ascontiguousarray(beta_array, dtype=np.float64)
    b0, b1 = (0, b.size) if rows is None else (int(rows[0]), int(rows[1]))
    p = np.empty((b1 - b0, t.size))
    norm = np.empty((b1 - b0, t.size))
    prob = problem.c_struct()
    _lib.check(lib.qw_rk4_grid_host(ctypes.byref(prob), t.ctypes.data, t.size, b.ctypes.data,
                                    b.size, b0, b1, p.ctypes.data, norm.ctypes.data,
                                    int(device)))
    return p, norm


def rk4_batch_host(problem, T, gamma, n_steps=None, return_psi=False, device=0):
    lib = _lib.load()
    T = np.ascontiguousarray(np.atleast_1d(T), dtype=np.float64)
    g = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(gamma), T.shape), dtype=np.float64)
    S = None if n_steps is None else np.ascontiguousarray(
        np.broadcast_to(np.asarray(n_steps), T.shape), dtype=np.int64)
    p, norm = np.empty(T.size), np.empty(T.size)
    psi = np.empty((T.size, int(problem.n)), dtype=np.complex128) if return_psi else None
    prob = problem.c_struct(per_point_steps=n_steps is not None)
    _lib.check(lib.qw_rk4_batch_host(
        ctypes.byref(prob), T.ctypes.data, g.ctypes.data, None if S is None else S.ctypes.data,
        T.size, p.ctypes.data, norm.ctypes.data, None if psi is None else psi.ctypes.data,
        int(device)))
    return (p, norm, psi) if return_psi else (p, norm)


def rk45_batch(problem, T, gamma, rtol=1e-6, atol=1e-6, order=None, return_psi=False,
               return_nfev=False, device=None):
    """The reference's shipped integrator -- SciPy RK45 with its step-size controller
    (BCB.py:177-190) -- on the device.  Returns (p, norm[, psi][, nfev])."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        dT, host = _as_dev(T, dev)
        dG, _ = _as_dev(gamma, dev)
        if dG.numel() == 1 and dT.numel() > 1:
            dG = dG.expand(dT.numel()).contiguous()
        if dT.numel() != dG.numel():
            raise ValueError("T and gamma must have the same length")
        P = dT.numel()
        dO = _as_dev(order, dev, torch.int64)[0] if order is not None else None
        p = torch.empty(P, dtype=torch.float64, device=dev)
        norm = torch.empty(P, dtype=torch.float64, device=dev)
        psi = (torch.empty((P, int(problem.n)), dtype=torch.complex128, device=dev)
               if return_psi else None)
        nfev = torch.zeros(P, dtype=torch.int64, device=dev) if return_nfev else None
        prob = problem.c_struct(per_point_steps=True)
        _lib.check(lib.qw_rk45_batch_dev(ctypes.byref(prob), _ptr(dT), _ptr(dG), _ptr(dO), P,
                                         float(rtol), float(atol), _ptr(p), _ptr(norm),
                                         _ptr(psi), _ptr(nfev), _stream()))
        out = (p, norm) + ((psi,) if return_psi else ()) + ((nfev,) if return_nfev else ())
        if host:
            out = tuple(t.cpu().numpy() for t in out)
        return out


def rk45_heatmap(problem, time_array, beta_array, rtol=1e-6, atol=1e-6, rows=None, device=None):
    """probability[beta][T] with the adaptive RK45 integrator of the reference."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        dT, host = _as_dev(time_array, dev)
        dB, _ = _as_dev(beta_array, dev)
        nT, nB = dT.numel(), dB.numel()
        b0, b1 = (0, nB) if rows is None else (int(rows[0]), int(rows[1]))
        p = torch.empty((b1 - b0, nT), dtype=torch.float64, device=dev)
        norm = torch.empty((b1 - b0, nT), dtype=torch.float64, device=dev)
        prob = problem.c_struct(per_point_steps=True)
        _lib.check(lib.qw_rk45_grid_dev(ctypes.byref(prob), _ptr(dT), nT, _ptr(dB), nB, b0, b1,
                                        float(rtol), float(atol), _ptr(p), _ptr(norm),
                                        _stream()))
        if host:
            return p.cpu().numpy(), norm.cpu().numpy()
        return p, norm


def ti_exact_batch(problem, time, gamma, order=None, return_psi=False, device=None):
    """|<w|exp(-iHt)|s>|^2 (not renormalised) and ||psi||^2 [, psi] for independent (t, gamma)
    points: the Chebyshev propagator behind evaluate_time_independent_probability
    (BCB.py:137-159).  Cycle graph, register-resident dimensions."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        dT, host = _as_dev(time, dev)
        dG, _ = _as_dev(gamma, dev)
        if dG.numel() == 1 and dT.numel() > 1:
            dG = dG.expand(dT.numel()).contiguous()
        if dT.numel() != dG.numel():
            raise ValueError("time and gamma must have the same length")
        P = dT.numel()
        dO = _as_dev(order, dev, torch.int64)[0] if order is not None else None
        p = torch.empty(P, dtype=torch.float64, device=dev)
        norm = torch.empty(P, dtype=torch.float64, device=dev)
        psi = (torch.empty((P, int(problem.n)), dtype=torch.complex128, device=dev)
               if return_psi else None)
        prob = problem.c_struct(per_point_steps=True)
        _lib.check(lib.qw_ti_exact_batch_dev(ctypes.byref(prob), _ptr(dT), _ptr(dG), _ptr(dO), P,
                                             _ptr(p), _ptr(norm), _ptr(psi), _stream()))
        out = (p, norm) + ((psi,) if return_psi else ())
        if host:
            out = tuple(t.cpu().numpy() for t in out)
        return out


def ti_exact_heatmap(problem, time_array, beta_array, rows=None, device=None):
    """Time-independent heatmap probability[gamma][t] with the exact propagator."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        dT, host = _as_dev(time_array, dev)
        dB, _ = _as_dev(beta_array, dev)
        nT, nB = dT.numel(), dB.numel()
        b0, b1 = (0, nB) if rows is None else (int(rows[0]), int(rows[1]))
        p = torch.empty((b1 - b0, nT), dtype=torch.float64, device=dev)
        norm = torch.empty((b1 - b0, nT), dtype=torch.float64, device=dev)
        prob = problem.c_struct(per_point_steps=True)
        _lib.check(lib.qw_ti_exact_grid_dev(ctypes.byref(prob), _ptr(dT), nT, _ptr(dB), nB, b0,
                                            b1, _ptr(p), _ptr(norm), _stream()))
        if host:
            return p.cpu().numpy(), norm.cpu().numpy()
        return p, norm


def rhs(problem, t, T, gamma, y, device=None):
    """-i H(t) y for one state: schrodinger_equation(t, y, beta, T) (BCB.py:163-174)."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        host = not (isinstance(y, torch.Tensor) and y.is_cuda)
        dy_in = (y if not host else torch.from_numpy(
            np.ascontiguousarray(np.asarray(y), dtype=np.complex128)).to(dev))
        dy_in = dy_in.to(torch.complex128).contiguous()
        if dy_in.numel() != int(problem.n):
            raise ValueError("state has %d entries, dimension is %d" % (dy_in.numel(), problem.n))
        out = torch.empty_like(dy_in)
        prob = problem.c_struct(per_point_steps=True)
        _lib.check(lib.qw_rhs_dev(ctypes.byref(prob), float(t), float(T), float(gamma),
                                  _ptr(dy_in), _ptr(out), _stream()))
        return out.cpu().numpy() if host else out


def adiabatic_batch(problem, s, gamma, reference_derivative=False, device=None):
    """Gap E1 - E0, coupling |<phi1|dH/ds|phi0>| and E0 of H(s) for (s_q, gamma_q) points:
    compute_energy_diff / compute_gamma of Schroedinger_Solver.py:208-253, batched on the
    device (secular equation of the rank-one perturbation; no dense eigensolve)."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        ds, host = _as_dev(s, dev)
        dg, _ = _as_dev(gamma, dev)
        if dg.numel() == 1 and ds.numel() > 1:
            dg = dg.expand(ds.numel()).contiguous()
        if ds.numel() != dg.numel():
            raise ValueError("s and gamma must have the same length")
        P = ds.numel()
        gap = torch.empty(P, dtype=torch.float64, device=dev)
        cpl = torch.empty(P, dtype=torch.float64, device=dev)
        e0 = torch.empty(P, dtype=torch.float64, device=dev)
        prob = problem.c_struct(per_point_steps=True)
        _lib.check(lib.qw_adiabatic_batch_dev(ctypes.byref(prob), _ptr(ds), _ptr(dg), P,
                                              1 if reference_derivative else 0, _ptr(gap),
                                              _ptr(cpl), _ptr(e0), _stream()))
        out = (gap, cpl, e0)
        if host:
            out = tuple(t.cpu().numpy() for t in out)
        return out


def spectrum_batch(problem, s, gamma, device=None):
    """All eigenvalues of H(s_q), ascending, for (s_q, gamma_q) points -> float64[P, n]:
    compute_eigenvalues of Eigenvalues_Crossing.py:59-71 (s = time / TIME), batched on the
    device (every root of the secular equation; no dense eigensolve)."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        ds, host = _as_dev(s, dev)
        dg, _ = _as_dev(gamma, dev)
        if dg.numel() == 1 and ds.numel() > 1:
            dg = dg.expand(ds.numel()).contiguous()
        if ds.numel() != dg.numel():
            raise ValueError("s and gamma must have the same length")
        P = ds.numel()
        eig = torch.empty((P, int(problem.n)), dtype=torch.float64, device=dev)
        prob = problem.c_struct(per_point_steps=True)
        _lib.check(lib.qw_spectrum_batch_dev(ctypes.byref(prob), _ptr(ds), _ptr(dg), P, _ptr(eig),
                                             _stream()))
        return eig.cpu().numpy() if host else eig


def adiabatic_check(problem, gamma, samples=4097, rounds=4, reference_derivative=False,
                    device=None):
    """max_s |<phi1|dH/ds|phi0>| and min_s (E1 - E0) over s in (0, 1), the two numbers behind
    adiabatic_theorem_check (Schroedinger_Solver.py:255-280: two shgo runs over dense
    eigensolves there): one launch over a uniform grid of interior points, then `rounds`
    launches that zoom into the best cell of each quantity (x 64 per round).  Returns
    dict(gamma_max, s_gamma, energy_min, s_gap, adiabatic_time)."""
    K = int(samples)
    s = (np.arange(K) + 0.5) / K
    gap, cpl, _ = adiabatic_batch(problem, s, np.full(K, float(gamma)),
                                  reference_derivative=reference_derivative, device=device)
    out = {}
    for name, vals, best in (("gamma", cpl, np.nanargmax), ("gap", gap, np.nanargmin)):
        i = int(best(vals))
        lo, hi = max(s[i] - 1.0 / K, 0.0), min(s[i] + 1.0 / K, 1.0)
        sb, vb = s[i], vals[i]
        for _ in range(int(rounds)):
            ss = np.linspace(lo, hi, 131)[1:-1]
            g2, c2, _ = adiabatic_batch(problem, ss, np.full(ss.size, float(gamma)),
                                        reference_derivative=reference_derivative,
                                        device=device)
            v2 = c2 if name == "gamma" else g2
            j = int(best(v2))
            sb, vb = ss[j], v2[j]
            w = (hi - lo) / 130.0
            lo, hi = max(sb - w, 0.0), min(sb + w, 1.0)
        out[name] = (float(vb), float(sb))
    gmax, emin = out["gamma"][0], out["gap"][0]
    return {"gamma_max": gmax, "s_gamma": out["gamma"][1], "energy_min": emin,
            "s_gap": out["gap"][1], "adiabatic_time": gmax / (emin * emin)}


def ensemble_stats(p, thresholds=(0.8, 0.9, 0.95, 0.99), device=None):
    """mean / std / min / max / fraction >= threshold over an ensemble of probabilities
    (contour levels of SS:415; thesis section 2.3)."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        dp, _ = _as_dev(np.asarray(p).ravel() if not isinstance(p, torch.Tensor)
                        else p.reshape(-1), dev)
        thr, _ = _as_dev(np.asarray(thresholds, dtype=np.float64), dev)
        out = torch.empty(4 + thr.numel(), dtype=torch.float64, device=dev)
        _lib.check(lib.qw_ensemble_stats_dev(_ptr(dp), dp.numel(), _ptr(thr), thr.numel(),
                                             _ptr(out), _stream()))
        o = out.cpu().numpy()
    return {"mean": float(o[0]), "std": float(o[1]), "min": float(o[2]), "max": float(o[3]),
            "fraction_ge": {float(t): float(f) for t, f in zip(thresholds, o[4:])}}


def robustness_map(probability, variation, device=None):
    """R[beta][T] of Robustness.py:25-31 (unrounded) and the per-column first-max beta index
    (Robustness.py:41-45)."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        host = not (isinstance(probability, torch.Tensor) and probability.is_cuda)
        dp = (probability if not host else torch.from_numpy(
            np.ascontiguousarray(probability, dtype=np.float64)).to(dev)).contiguous()
        nb, nt = dp.shape
        R = torch.empty_like(dp)
        arg = torch.empty(nt, dtype=torch.int64, device=dev)
        _lib.check(lib.qw_heatmap_robustness_dev(_ptr(dp), nb, nt, int(variation), _ptr(R),
                                                 _ptr(arg), _stream()))
        if host:
            return R.cpu().numpy(), arg.cpu().numpy()
        return R, arg


def heatmap_summary(probability, time_array, t_min=0.0, device=None):
    """Best cell of a heatmap and the thesis' figures of merit (Eq. 2.10-2.11): first-maximum
    (beta, T) indices of p, I = 1/p_max, and tau = min T/p over T >= t_min."""
    dev = _device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        dp, _ = _as_dev(probability if isinstance(probability, torch.Tensor)
                        else np.ascontiguousarray(probability, dtype=np.float64).ravel(), dev)
        nb, nt = probability.shape
        dt, _ = _as_dev(time_array, dev)
        out = torch.empty(6, dtype=torch.float64, device=dev)
        _lib.check(lib.qw_heatmap_summary_dev(_ptr(dp), _ptr(dt), nb, nt, float(t_min),
                                              _ptr(out), _stream()))
        o = out.cpu().numpy()
    return {"p_max": float(o[0]), "argmax": (int(o[1]), int(o[2])),
            "iterations": float(1.0 / o[0]) if o[0] > 0 else float("inf"),
            "tau_min": float(o[3]), "tau_argmin": (int(o[4]), int(o[5]))}


def _halton(n, base):
    out = np.zeros(n)
    for i in range(n):
        f, k, x = 1.0, i + 1, 0.0
        while k > 0:
            f /= base
            x += f * (k % base)
            k //= base
        out[i] = x
    return out


def maximize_probability(problem, beta_bounds, time_bounds, n_samples=4096, keep=16, rounds=14,
                         stencil=5, device=None):
    """Batched maximisation of p(beta, T) -- the job of the reference's SciPy optimisers
    (Schroedinger_Solver.py:283-361: shgo / basinhopping / dual_annealing, one point per
    objective call) done with whole candidate sets per kernel launch.

    One launch scans a low-discrepancy (Halton) sample of the box; the ``keep`` best cells
    are then refined together by a shrinking ``stencil`` x ``stencil`` pattern search, one
    launch per round.  Returns (p_max, beta, T, local_maxima[keep, 3]) with rows
    (p, beta, T) sorted by p, like ``arrange_local_maxima`` (SS:134-144)."""
    dev = _device(device)
    b0, b1 = float(beta_bounds[0]), float(beta_bounds[1])
    t0, t1 = float(time_bounds[0]), float(time_bounds[1])
    with torch.cuda.device(dev):
        beta = torch.from_numpy(b0 + (b1 - b0) * _halton(n_samples, 2)).to(dev)
        time = torch.from_numpy(t0 + (t1 - t0) * _halton(n_samples, 3)).to(dev)
        p, _ = rk4_batch(problem, time, beta, device=dev)
        top = torch.topk(p, min(keep, n_samples))
        cb, ct, cp = beta[top.indices].clone(), time[top.indices].clone(), top.values.clone()
        k = cb.numel()
        rb = torch.full((k,), (b1 - b0) / np.sqrt(n_samples), dtype=torch.float64, device=dev)
        rt = torch.full((k,), (t1 - t0) / np.sqrt(n_samples), dtype=torch.float64, device=dev)
        off = torch.linspace(-1.0, 1.0, stencil, dtype=torch.float64, device=dev)
        ob, ot = torch.meshgrid(off, off, indexing="ij")
        ob, ot = ob.reshape(1, -1), ot.reshape(1, -1)
        for _ in range(rounds):
            gb = (cb[:, None] + rb[:, None] * ob).clamp(b0, b1)
            gt = (ct[:, None] + rt[:, None] * ot).clamp(t0, t1)
            gp, _ = rk4_batch(problem, gt.reshape(-1), gb.reshape(-1), device=dev)
            gp = gp.reshape(k, -1)
            best = gp.argmax(dim=1)
            bp = gp.gather(1, best[:, None])[:, 0]
            improved = bp > cp
            cb = torch.where(improved, gb.gather(1, best[:, None])[:, 0], cb)
            ct = torch.where(improved, gt.gather(1, best[:, None])[:, 0], ct)
            cp = torch.where(improved, bp, cp)
            shrink = torch.where(improved, 1.0, 0.5)
            rb, rt = rb * shrink, rt * shrink
        order = torch.argsort(cp, descending=True)
        loc = torch.stack([cp[order], cb[order], ct[order]], dim=1).cpu().numpy()
    return float(loc[0, 0]), float(loc[0, 1]), float(loc[0, 2]), loc


def plan(problem, device=None):
    dev = _device(device)
    lib = _lib.load()
    pl = _lib.QwPlan()
    with torch.cuda.device(dev):
        _lib.check(lib.qw_plan_query(ctypes.byref(problem.c_struct()), ctypes.byref(pl)))
    d = {k: getattr(pl, k) for k, _ in _lib.QwPlan._fields_}
    d["path"] = _lib.PATH_NAMES.get(pl.path, str(pl.path))
    return d


def fp64_peak_tflops(millis=200.0, device=None):
    dev = _device(device)
    lib = _lib.load()
    out = ctypes.c_double(0.0)
    with torch.cuda.device(dev):
        _lib.check(lib.qw_fp64_peak_probe(float(millis), ctypes.byref(out), _stream()))
    return out.value


def launch_count():
    return int(_lib.load().qw_launch_count())


# ---- sharding over GPUs: one process per GPU, rows of the beta axis ------------------------

def shard_rows(n_rows, world_size, rank):
    """Contiguous beta-row blocks, like the Ray tasks of BCB.py:279-281; uneven row counts
    are spread over the first ranks."""
    base, extra = divmod(int(n_rows), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def heatmap_sharded(problem, time_array, beta_array, group=None, compute=None,
                    gather=True, dst=None):
    """Every rank integrates its block of beta rows; ONE collective assembles
    probability[beta][T] (``ray.get`` + ``np.concatenate`` of BCB.py:284-292).

    ``compute(problem, time_array, beta_array, rows) -> (p, norm)`` defaults to the CUDA
    path; it may return CUDA tensors (they stay on the device through the collective),
    CPU tensors or numpy arrays (CPU tests of the sharding logic inject the oracle here).
    ``dst=None``: every rank gets the full heatmap (all_gather_into_tensor + one D2H per
    rank); ``dst=r``: only rank r copies it to the host, the others return ``(None, None)``
    (the reference's driver process is the only reader).  ``gather=False``: this rank's
    block.  Returns numpy arrays (p, norm).
    """
    import torch.distributed as dist
    time_array = np.asarray(time_array, dtype=np.float64)
    beta_array = np.asarray(beta_array, dtype=np.float64)
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        backend = dist.get_backend(group)
    else:
        world, rank, backend = 1, 0, None
    b0, b1 = shard_rows(beta_array.size, world, rank)
    if compute is None:
        dev = _device()
        p, norm = rk4_heatmap(problem, torch.from_numpy(time_array).to(dev),
                              torch.from_numpy(beta_array).to(dev), rows=(b0, b1))
    else:
        p, norm = compute(problem, time_array, beta_array, (b0, b1))
        if not isinstance(p, torch.Tensor):
            p = torch.from_numpy(np.ascontiguousarray(p))
            norm = torch.from_numpy(np.ascontiguousarray(norm))
    if world == 1 or not gather:
        return p.cpu().numpy(), norm.cpu().numpy()
    if backend == "nccl" and not p.is_cuda:          # NCCL moves device memory only
        dev = _device()
        p, norm = p.to(dev), norm.to(dev)
    # equal blocks (padded to the largest): one all_gather_into_tensor, one D2H
    nT = time_array.size
    rows_max = -(-beta_array.size // world)
    send = torch.zeros((2, rows_max, nT), dtype=torch.float64, device=p.device)
    send[0, :b1 - b0] = p
    send[1, :b1 - b0] = norm
    recv = torch.empty((world * 2, rows_max, nT), dtype=torch.float64, device=p.device)
    dist.all_gather_into_tensor(recv, send, group=group)     # rank blocks along dim 0
    if dst is not None and rank != dst:
        return None, None
    host = recv.cpu().numpy().reshape(world, 2, rows_max, nT)    # the single D2H copy
    counts = [shard_rows(beta_array.size, world, r) for r in range(world)]
    ps = np.concatenate([host[r, 0, :r1 - r0] for r, (r0, r1) in enumerate(counts)])
    ns = np.concatenate([host[r, 1, :r1 - r0] for r, (r0, r1) in enumerate(counts)])
    return ps, ns


def sharded_d2h_bytes(n_beta, n_time, world, rank, dst=None):
    """Bytes ``heatmap_sharded`` copies device -> host on ``rank`` (bench.py's e2e account)."""
    if world == 1:
        return 2 * n_beta * n_time * 8
    if dst is not None and rank != dst:
        return 0
    return world * 2 * (-(-n_beta // world)) * n_time * 8
