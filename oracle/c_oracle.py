"""ctypes front-end of oracle/rk4_oracle.c.  TEST INFRASTRUCTURE ONLY."""

import ctypes
import os
import subprocess

import numpy as np

from .defs import (schedule_id, topology_id, gamma_on_id, default_target)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librk4_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "rk4_oracle.c")
    if (not force and os.path.isfile(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                   stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        _lib.qwo_rk4_batch.restype = ctypes.c_int
        _lib.qwo_rk4_batch.argtypes = [
            ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
            dp, dp, ip, ctypes.c_double, ctypes.c_int64, dp, dp, dp, ctypes.c_int]
        _lib.qwo_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    """Host threads this process may use.  Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1, which would silently make the host baseline single-threaded."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def rk4_batch(n, topology, schedule, T, gamma, n_steps=None, step_size=None,
              gamma_on="oracle", target=None, return_psi=False, n_threads=0):
    """(p[P], norm[P][, psi[P,N]], threads_used) for independent points."""
    T = np.ascontiguousarray(np.atleast_1d(T), dtype=np.float64)
    gamma = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(gamma), T.shape),
                                 dtype=np.float64)
    P = T.shape[0]
    if step_size is not None:
        h = float(step_size)
        S = np.array([int(t / h) for t in T], dtype=np.int64)
    else:
        h = 0.0
        S = np.ascontiguousarray(np.broadcast_to(np.asarray(n_steps, dtype=np.int64),
                                                 (P,)))
    w = default_target(n) if target is None else int(target)
    p = np.empty(P)
    norm = np.empty(P)
    psi = np.empty((P, n), dtype=np.complex128) if return_psi else None
    dp = ctypes.POINTER(ctypes.c_double)
    rc = lib().qwo_rk4_batch(
        int(n), topology_id(topology), schedule_id(schedule), gamma_on_id(gamma_on), w,
        T.ctypes.data_as(dp), gamma.ctypes.data_as(dp),
        S.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), h, P,
        p.ctypes.data_as(dp), norm.ctypes.data_as(dp),
        psi.ctypes.data_as(dp) if return_psi else None,
        int(n_threads) if n_threads > 0 else max_threads())
    if rc < 0:
        raise ValueError("qwo_rk4_batch failed with %d" % rc)
    if return_psi:
        return p, norm, psi, rc
    return p, norm, rc


def heatmap(n, topology, schedule, time_array, beta_array, **kw):
    """p[n_beta, n_T], norm[n_beta, n_T] (layout of BCB.py:239-259)."""
    time_array = np.asarray(time_array, dtype=np.float64)
    beta_array = np.asarray(beta_array, dtype=np.float64)
    bb, tt = np.meshgrid(beta_array, time_array, indexing="ij")
    out = rk4_batch(n, topology, schedule, tt.ravel(), bb.ravel(), **kw)
    return out[0].reshape(bb.shape), out[1].reshape(bb.shape)
