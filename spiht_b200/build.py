"""Builds libspiht_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The built library lives next to this file so that it travels with the source
tree; nothing is installed into site-packages.
"""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libspiht_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
    "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(_HERE, "..", "include", "spiht_b200.h"), os.path.abspath(__file__)]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libspiht_b200.so cannot be built")
    return nvcc


def build(force=False, verbose=False, extra_flags=()):
    """Compile every CUDA source into libspiht_b200.so; returns the path."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-o", LIB_PATH] + sources()
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose and proc.stdout:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True,
                extra_flags=("-Xptxas", "-v") if "--ptxas" in sys.argv else ()))
