"""Builds liblossyless_b200.so (sm_100a) in-tree with nvcc.

    python -m lossyless_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/lossyless_b200.h); it links the CUDA
runtime statically and resolves the one driver entry point it needs
(cuTensorMapEncodeTiled) at run time, so it builds on a machine without a GPU driver.
"""
import argparse
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "liblossyless_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(os.path.dirname(HERE), "include", "lossyless_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; liblossyless_b200.so cannot be built")
    return nvcc


def build(force=False, verbose=False, jobs=None):
    """Compile every .cu in csrc/ to an object (in parallel) and link the shared library."""
    if not force and not _stale():
        return LIB
    nvcc = find_nvcc()
    os.makedirs(OUT_DIR, exist_ok=True)
    obj_dir = os.path.join(OUT_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        hdr_t = max(os.path.getmtime(h) for h in glob.glob(os.path.join(CSRC, "*.cuh")) + [
            os.path.join(os.path.dirname(HERE), "include", "lossyless_b200.h")])
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            procs.append((src, obj, None))
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        if p is not None:
            out, _ = p.communicate()
            if p.returncode != 0:
                sys.stderr.write(out)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                sys.stderr.write(out)
        objs.append(obj)
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of liblossyless_b200.so failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
