"""Builds libnbody_b200.so (hand-written sm_100a CUDA kernels + the C ABI) and the host executable, in-tree.

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libnbody_b200.so")
EXE = os.path.join(HERE, "N_Body_Simulation")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-ccbin", HOST_CXX, "--expt-relaxed-constexpr"]

# developer A/B builds: extra nvcc flags (e.g. -DNB_NODE_WORDS=8) for a library built into another directory
NVCC_FLAGS += os.environ.get("NB_NVCC_EXTRA", "").split()

CU_SOURCES = ["api.cu", "naive.cu", "integrator.cu", "energy.cu", "bh_build.cu", "bh_traverse.cu", "comm.cu", "util.cu"]
HOST_SOURCES = ["main.cpp", "Configuration.cpp", "InputParser.cpp", "StateFile.cpp", "TimeConverter.cpp", "TimeMeasurement.cpp",
                "nBodyAlgorithm.cpp", "NaiveAlgorithm.cpp", "BarnesHutAlgorithm.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_library(force=False, verbose=False, ptxas_info=False, out_dir=None):
    """out_dir: build objects and the library there instead of in-tree (a clean rebuild that leaves the tree alone)."""
    objdir = os.path.join(out_dir or HERE, "build")
    lib = os.path.join(out_dir, "libnbody_b200.so") if out_dir else LIB
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "nbody_b200.h"))
    objs = []
    procs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
    failed = [src for src, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for: " + ", ".join(failed))
    if force or procs or _newer(lib, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs +
             ["-lcudart", "-ldl", "-Xlinker", "-z,defs"], verbose)
    return lib


def build_host(force=False, verbose=False):
    """The reference-facing executable (same flags as the reference's main.cpp) on top of the C ABI."""
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES]
    if not all(os.path.exists(s) for s in srcs):
        return None
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")] + [os.path.join(ROOT, "include", "nbody_b200.h")]
    if force or _newer(EXE, srcs + hdrs + [LIB]):
        _run([HOST_CXX, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST] + srcs +
             ["-o", EXE, "-L", HERE, "-lnbody_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + HERE], verbose)
    return EXE


def build_all(force=False, verbose=False):
    lib = build_library(force, verbose)
    exe = build_host(force, verbose)
    return lib, exe


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
