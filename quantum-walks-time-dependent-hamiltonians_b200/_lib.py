"""ctypes binding of libqw_b200.so (include/qw_b200.h).  No CPU fallback: if the CUDA
library is missing or a call fails, the caller gets an exception."""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QW_B200_LIB: another build of the same library (A/B timing of two versions, tools/r2_ab_sizes.sh)
LIB_PATH = os.environ.get("QW_B200_LIB") or os.path.join(_HERE, "libqw_b200.so")

# enums of include/qw_b200.h
TOPO_COMPLETE, TOPO_CYCLE = 0, 1
SCHEDULES = {"cerf_3": 0, "lin": 1, "sqrt": 2, "cbrt": 3, "cerf_5": 4, "cerf_7": 5,
             "static": 6}
GAMMA_ON_ORACLE, GAMMA_ON_LAPLACIAN = 0, 1
FLAG_FORCE_GENERIC, FLAG_FORCE_STREAM, FLAG_EXACT_SUM = 1, 2, 4
FLAG_STREAM_KB1, FLAG_STREAM_KB2 = 8, 16
FLAG_NO_HORNER, FLAG_MIRROR = 2048, 8192
PATH_NAMES = {0: "resident-cycle", 1: "resident-complete", 2: "generic", 3: "stream",
              4: "packed-cycle", 5: "horner-cycle", 6: "horner-complete",
              7: "packed-horner-cycle", 8: "packed-any-cycle", 9: "mirror-cycle",
              10: "symmetric-complete", 11: "packed-complete", 12: "cluster-cycle",
              13: "packed-mirror-cycle", 14: "parallel-in-time"}


class QwProblem(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("topology", ctypes.c_int32),
                ("schedule", ctypes.c_int32), ("gamma_on", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("target", ctypes.c_int64),
                ("step_size", ctypes.c_double), ("n_steps", ctypes.c_int64)]


class QwGridJob(ctypes.Structure):
    _fields_ = [("prob", QwProblem), ("time_array", ctypes.c_void_p), ("n_time", ctypes.c_int64),
                ("beta_array", ctypes.c_void_p), ("n_beta", ctypes.c_int64),
                ("beta_begin", ctypes.c_int64), ("beta_end", ctypes.c_int64),
                ("p", ctypes.c_void_p), ("norm", ctypes.c_void_p)]


class QwPlan(ctypes.Structure):
    _fields_ = [("path", ctypes.c_int32), ("sites_per_thread", ctypes.c_int32),
                ("threads_per_block", ctypes.c_int32), ("blocks_per_sm", ctypes.c_int32),
                ("regs_per_thread", ctypes.c_int32), ("smem_bytes", ctypes.c_int32),
                ("sm_count", ctypes.c_int32), ("steps_per_pass", ctypes.c_int32)]


_vp, _i64, _i32, _dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double
_pp = ctypes.POINTER(QwProblem)

# name -> (restype, argtypes); tests/test_abi.py checks this table against the header
SIGNATURES = {
    "qw_rk4_batch_dev": (ctypes.c_int, [_pp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "qw_rk4_grid_dev": (ctypes.c_int, [_pp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "qw_rk4_multi_grid_dev": (ctypes.c_int, [ctypes.POINTER(QwGridJob), _i64, _vp]),
    "qw_rk4_batch_host": (ctypes.c_int, [_pp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, ctypes.c_int]),
    "qw_rk4_grid_host": (ctypes.c_int, [_pp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp,
                                        ctypes.c_int]),
    "qw_ti_exact_batch_dev": (ctypes.c_int, [_pp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "qw_ti_exact_grid_dev": (ctypes.c_int, [_pp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "qw_rk45_batch_dev": (ctypes.c_int, [_pp, _vp, _vp, _vp, _i64, _dbl, _dbl, _vp, _vp, _vp, _vp,
                                         _vp]),
    "qw_rk45_grid_dev": (ctypes.c_int, [_pp, _vp, _i64, _vp, _i64, _i64, _i64, _dbl, _dbl, _vp,
                                        _vp, _vp]),
    "qw_rhs_dev": (ctypes.c_int, [_pp, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
    "qw_ensemble_stats_dev": (ctypes.c_int, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "qw_heatmap_robustness_dev": (ctypes.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "qw_heatmap_summary_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _dbl, _vp, _vp]),
    "qw_adiabatic_batch_dev": (ctypes.c_int, [_pp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "qw_spectrum_batch_dev": (ctypes.c_int, [_pp, _vp, _vp, _i64, _vp, _vp]),
    "qw_scratch_trim": (ctypes.c_int, []),
    "qw_plan_query": (ctypes.c_int, [_pp, ctypes.POINTER(QwPlan)]),
    "qw_fp64_peak_probe": (ctypes.c_int, [_dbl, ctypes.POINTER(ctypes.c_double), _vp]),
    "qw_launch_count": (ctypes.c_int64, []),
    "qw_last_error": (ctypes.c_char_p, []),
    "qw_version": (ctypes.c_int, []),
}

_lib = None


def load():
    """Load the CUDA library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    """Map the C return convention onto exceptions."""
    if rc == 0:
        return
    msg = load().qw_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError("qw_b200: %s (code %d)" % (msg, rc))
    raise RuntimeError("qw_b200: CUDA error %d: %s" % (rc, msg))
