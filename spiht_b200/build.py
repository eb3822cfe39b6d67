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
# SPIHTB_VARIANT=prof selects the debug build with the coders' phase counters (-DSPIHTB_PROF):
# libspiht_b200_prof.so, objects under csrc/_obj_prof (tools/dec_phases.py); the product library is the default.
VARIANT = os.environ.get("SPIHTB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "libspiht_b200%s.so" % ("_" + VARIANT if VARIANT else ""))

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
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
    # one nvcc process per translation unit (in parallel), objects under csrc/_obj, then one link
    nvcc = find_nvcc()
    objdir = os.path.join(CSRC, "_obj" + ("_" + VARIANT if VARIANT else ""))
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))
    flag_tag = os.path.join(objdir, "flags.txt")
    flags = NVCC_FLAGS + list(extra_flags) + (["-DSPIHTB_PROF"] if VARIANT == "prof" else []) + \
        (["-DSPIHTB_NO_LAZY"] if VARIANT == "nolazy" else [])   # A-B build without the lazy zero fill in the decoder
    same_flags = os.path.exists(flag_tag) and open(flag_tag).read() == " ".join(flags)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        fresh = (not force and same_flags and os.path.exists(obj)
                 and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t))
        if not fresh:
            cmd = [nvcc] + flags + ["-c", "-o", obj, src]
            if verbose:
                print(" ".join(cmd))
            jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = ""
    failed = False
    for cmd, proc in jobs:
        out, _ = proc.communicate()
        log += out or ""
        failed = failed or proc.returncode != 0
    if verbose and log:
        print(log)
    if failed:
        raise RuntimeError("nvcc failed:\n" + log)
    with open(flag_tag, "w") as f:
        f.write(" ".join(flags))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    proc = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    extra = ["-Xptxas", "-v"] if "--ptxas" in sys.argv else []
    extra += [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose=True, extra_flags=tuple(extra)))
