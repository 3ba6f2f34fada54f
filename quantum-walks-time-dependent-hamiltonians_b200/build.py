"""Compile csrc/*.cu into libqw_b200.so (in-tree, sm_100a only)."""

import concurrent.futures
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["qw_api.cu", "qw_resident.cu", "qw_generic.cu", "qw_stream.cu", "qw_packed.cu",
           "qw_cheb.cu", "qw_rk45.cu", "qw_aux.cu", "qw_horner.cu", "qw_adiabatic.cu",
           "qw_symmetric.cu", "qw_pit.cu"]
OUT = os.path.join(_HERE, "libqw_b200.so")
OBJ_DIR = os.path.join(_HERE, "build")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = glob.glob(os.path.join(_HERE, "csrc", "*")) + [
        os.path.join(os.path.dirname(_HERE), "include", "qw_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _env():
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports CC/CXX wrappers that nvcc must not pick up
    env.pop("CXX", None)
    return env


def build(force=False, verbose=False, extra_flags=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> libqw_b200.so.
    One nvcc per source file, in parallel, then one link step.  Returns the path."""
    if not force and not _stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = glob.glob(os.path.join(_HERE, "csrc", "*.cuh")) + [
        os.path.join(os.path.dirname(_HERE), "include", "qw_b200.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        path = os.path.join(_HERE, "csrc", src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if (not force and os.path.isfile(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(path), newest_header)):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", "-o", obj, path]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, env=_env())
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, env=_env())
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
