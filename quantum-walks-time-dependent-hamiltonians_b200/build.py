"""Compile csrc/*.cu into libqw_b200.so (in-tree, sm_100a only)."""

import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["qw_api.cu", "qw_resident.cu", "qw_generic.cu", "qw_stream.cu", "qw_packed.cu", "qw_cheb.cu",
           "qw_aux.cu"]
OUT = os.path.join(_HERE, "libqw_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = glob.glob(os.path.join(_HERE, "csrc", "*")) + [
        os.path.join(os.path.dirname(_HERE), "include", "qw_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libqw_b200.so.  Returns the path."""
    if not force and not _stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-o", OUT] + [
        os.path.join(_HERE, "csrc", s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports CC/CXX wrappers that nvcc must not pick up
    env.pop("CXX", None)
    subprocess.run(cmd, check=True, env=env)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
