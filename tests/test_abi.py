"""CPU: the C-ABI library loads and exports every symbol include/qw_b200.h declares, the
Python enum tables agree with the header and with the oracle, and argument errors are
reported through the return code without touching a GPU."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qw_b200.h")


@pytest.fixture(scope="module")
def lib(qw):
    import __graft_entry__ as g
    g.build()
    return qw._lib.load()


def _header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qw_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(qw, lib):
    names = _header_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(qw._lib.SIGNATURES) == names
    assert lib.qw_version() == 1


def test_enum_tables_agree(qw):
    text = open(HEADER).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(QW_[A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    from oracle import defs
    L = qw._lib
    assert enums["QW_TOPO_COMPLETE"] == L.TOPO_COMPLETE == defs.TOPOLOGY_COMPLETE == 0
    assert enums["QW_TOPO_CYCLE"] == L.TOPO_CYCLE == defs.TOPOLOGY_CYCLE == 1
    for name, hdr in (("cerf_3", "CERF3"), ("lin", "LIN"), ("sqrt", "SQRT"), ("cbrt", "CBRT"),
                      ("cerf_5", "CERF5"), ("cerf_7", "CERF7"), ("static", "STATIC")):
        assert enums["QW_SCHED_" + hdr] == L.SCHEDULES[name] == defs.SCHEDULES[name]
    assert enums["QW_GAMMA_ON_LAPLACIAN"] == L.GAMMA_ON_LAPLACIAN == defs.GAMMA_ON_LAPLACIAN
    assert enums["QW_FLAG_FORCE_GENERIC"] == L.FLAG_FORCE_GENERIC
    assert enums["QW_FLAG_FORCE_STREAM"] == L.FLAG_FORCE_STREAM
    assert enums["QW_FLAG_EXACT_SUM"] == L.FLAG_EXACT_SUM


def test_struct_layouts_match_header(qw):
    assert ctypes.sizeof(qw._lib.QwProblem) == 48
    assert ctypes.sizeof(qw._lib.QwPlan) == 32


def test_argument_errors_need_no_gpu(qw, lib):
    L = qw._lib
    bad = [
        (L.QwProblem(n=2, topology=1, schedule=1, n_steps=4), -2),      # cycle N < 3
        (L.QwProblem(n=8, topology=7, schedule=1, n_steps=4), -3),      # topology
        (L.QwProblem(n=8, topology=1, schedule=9, n_steps=4), -3),      # schedule
        (L.QwProblem(n=8, topology=1, schedule=1, gamma_on=5, n_steps=4), -3),
        (L.QwProblem(n=8, topology=1, schedule=1, n_steps=-1), -4),     # no step rule
        (L.QwProblem(n=8, topology=1, schedule=1, target=8, n_steps=1), -2),
    ]
    for prob, code in bad:
        rc = lib.qw_rk4_batch_dev(ctypes.byref(prob), None, None, None, None, 0, None, None,
                                  None, None)
        assert rc == code, (rc, code)
        assert lib.qw_last_error()
    ok = L.QwProblem(n=8, topology=1, schedule=1, n_steps=4)
    assert lib.qw_rk4_batch_dev(ctypes.byref(ok), None, None, None, None, 3, None, None,
                                None, None) == -1                        # NULL arrays
    assert lib.qw_rk4_batch_dev(None, None, None, None, None, 0, None, None, None, None) == -1
    assert lib.qw_rk4_batch_dev(ctypes.byref(ok), None, None, None, None, 0, None, None,
                                None, None) == 0                         # empty batch
    assert lib.qw_rk4_grid_dev(ctypes.byref(ok), None, 4, None, 4, 3, 2, None, None, None) == -5
    with pytest.raises(ValueError):
        L.check(-3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "quantum-walks-time-dependent-hamiltonians_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            src = open(os.path.join(dirpath, f), errors="replace").read() \
                if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
            assert not re.search(r"#include\s+[\"<][^\">]*oracle", src), f
            assert "rk4_oracle" not in src and "librk4" not in src, f


def test_missing_library_fails_loudly(qw, monkeypatch, tmp_path):
    L = qw._lib
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.load()
