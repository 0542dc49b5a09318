"""Build recipes for the native parts (in-tree, sm_100a only).

    python -m svinet_b200.build            # libsvi_ls.so (+ host CLI when its sources exist)

nvcc cross-compiles for sm_100a without a GPU.  Outputs go to svinet_b200/lib/ (git-ignored,
shipped to the GPU box by gpurun).
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libsvi_ls.so")
CLI = os.path.join(LIBDIR, "svinet")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "--fmad=true", "-Xcompiler", "-fPIC,-O2,-Wall",
              "-Xptxas", "-v" if os.environ.get("SVI_PTXAS_V") else "-O3"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_lib(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "svi_ls.cu"), os.path.join(CSRC, "svi_fa2.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(CSRC, "svi_common.h"), os.path.join(REPO, "include", "svi_ls.h"), os.path.join(REPO, "include", "svi_fa2.h")]
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and not _newer(LIB, deps):
        return LIB
    cmd = [NVCC] + ARCH + NVCC_FLAGS + ["-shared", "-o", LIB] + srcs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


def build_cli(force=False, verbose=False):
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cc")) if os.path.isdir(HOST) else []
    if not srcs:
        return None
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hh")]
    deps = srcs + hdrs + [os.path.join(REPO, "include", "svi_ls.h"), LIB]
    if not force and not _newer(CLI, deps):
        return CLI
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O3", "-g", "-Wall", "-I", os.path.join(REPO, "include"),
           "-o", CLI] + srcs + ["-L", LIBDIR, "-lsvi_ls", "-Wl,-rpath,$ORIGIN", "-lpthread"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return CLI


def build_all(force=False, verbose=False):
    lib = build_lib(force, verbose)
    cli = build_cli(force, verbose)
    return lib, cli


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
