"""Load the reference's functions *unmodified* from /root/reference (build container only).

The reference scripts cannot be imported as they stand: they import seaborn / matplotlib /
ray (absent) and run a sweep at import time (Schroedinger_Solver.py:463-467).  Following
SURVEY.md Appendix C we parse each file, keep only its top-level import and def nodes,
and exec those into a fresh module.  Nothing is copied out of the reference tree.

TEST INFRASTRUCTURE ONLY -- used by tests/golden/make_golden.py and by CPU tests that are
skipped when /root/reference is absent (e.g. on the GPU box).
"""

import ast
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QW_REFERENCE_ROOT", "/root/reference")

_STUBS = {
    "seaborn": {},
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "matplotlib.patches": {"Rectangle": object},
    "matplotlib.colors": {},
    "ray": {"remote": staticmethod(lambda f: f)},
    "NumbaLSODA": {"lsoda_sig": None, "lsoda": None},
}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Code", "BCB.py"))


def _install_stubs():
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v.__func__ if isinstance(v, staticmethod) else v)
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)


def load(script, **module_globals):
    """Return a module holding the defs of ``Code/<script>.py`` with globals preset.

    Example: ``load('BCB', dimension=16, step_function='lin', topology='circular')``.
    """
    if not reference_available():
        raise FileNotFoundError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    path = os.path.join(REFERENCE_ROOT, "Code", script + ".py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [n for n in tree.body
            if isinstance(n, (ast.Import, ast.ImportFrom, ast.FunctionDef))]
    mod = types.ModuleType("ref_" + script)
    mod.__file__ = path
    code = compile(ast.Module(body=keep, type_ignores=[]), path, "exec")
    exec(code, mod.__dict__)
    for k, v in module_globals.items():
        setattr(mod, k, v)
    return mod
