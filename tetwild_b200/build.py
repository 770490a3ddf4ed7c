"""Build recipe for libtetwild_gpu.so (in-tree, sm_100a only).

    python -m tetwild_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU. The shared object is written next to this file so that it travels with the
source tree (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtetwild_gpu.so")
SOURCES = ["ctx.cu", "multi.cu", "qsort.cu", "amips.cu", "mesh.cu", "surface.cu", "envelope.cu", "winding.cu", "winding_build.cu", "peaks.cu"]
HEADERS = ["common.cuh", "tw_math.cuh", "winding_math.cuh", "winding.cuh", "surface.cuh", "sampling.cuh", os.path.join("..", "..", "include", "tetwild_gpu.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    env = dict(os.environ)
    # the image exports CC/CXX wrappers without OpenMP specs; nvcc only needs a plain host g++
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    procs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [_nvcc(), "-ccbin", ccbin] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
    cmd = [_nvcc(), "-ccbin", ccbin, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
