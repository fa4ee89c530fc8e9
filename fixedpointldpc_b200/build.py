"""Compile libldpc_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldpc_b200.so")
SOURCES = ["ldpc_decoder.cu", "ldpc_encode.cu", "ldpc_multi.cu", "ldpc_decode_f64.cu", "ldpc_code.cpp"]
HEADERS = ["ldpc_kernels.cuh", "ldpc_code.hpp", os.path.join("..", "..", "include", "ldpc_capi.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


FACADE_DIR = os.path.join(CSRC, "facade")
INCLUDE_DIR = os.path.join(HERE, "..", "include")
FACADE_VARIANTS = {"wifi": 0, "a5": 1, "a24": 2, "c79": 3}


def wrapper_path(variant):
    return os.path.join(HERE, "ldpc_wrapper_" + variant)


def build_facade(force=False):
    """The reference-compatible console program (PerfTest/Wrapper drivers on the engine), one binary per
    compile-time code variant like the reference itself."""
    srcs = [os.path.join(FACADE_DIR, f) for f in ("PerfTest.cpp", "Wrapper.cpp")]
    deps = srcs + [os.path.join(INCLUDE_DIR, f) for f in ("ArrayLDPCMacro.h", "ArrayLDPC.h", "PerfTest.h", "ldpc_capi.h")] + [LIB]
    for name, variant in FACADE_VARIANTS.items():
        out = wrapper_path(name)
        if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
            continue
        cmd = ["g++", "-O2", "-std=c++17", "-DLDPC_CODE_VARIANT=%d" % variant, "-I", INCLUDE_DIR, "-o", out] + srcs + \
              ["-L", HERE, "-lldpc_b200", "-Wl,-rpath,$ORIGIN"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("g++ failed building " + out)


def build(force=False, verbose=False):
    lib = build_lib(force, verbose)
    build_facade(force)
    return lib


def build_lib(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    # NCCL is bound at run time (dlopen in ldpc_multi.cu); only its header is needed here
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libldpc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
