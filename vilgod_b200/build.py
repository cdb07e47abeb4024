"""Builds libvilgod_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the
repo snapshot to the GPU box).  ``python -m vilgod_b200.build [--force] [-v]``"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(OUT_DIR, "libvilgod_b200.so")             # fp16 GEMM operands (default)
LIB_PATH_BF16 = os.path.join(OUT_DIR, "libvilgod_b200_bf16.so")   # bf16 GEMM operands (-DVG_OPERAND_BF16)
SOURCES = ["api.cu", "canonicalise.cu", "projection.cu", "gemm_tcgen05.cu", "attention_tcgen05.cu",
           "vit_misc.cu"]
HEADERS = ["common.cuh", "ptx.cuh", os.path.join("..", "..", "include", "vilgod_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


LIB_PATH_TRACE = os.path.join(OUT_DIR, "libvilgod_b200_trace.so")  # debug: -DVG_PROJ_TRACE (load with VG_LIB_PATH)


def build_library(force=False, verbose=False, bf16=False, trace=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, jobs = [], []
    lib_path = LIB_PATH_TRACE if trace else LIB_PATH_BF16 if bf16 else LIB_PATH
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, s.replace(".cu", "_trace.o" if trace else "_bf16.o" if bf16 else ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = ([_nvcc()] + NVCC_FLAGS + (["-DVG_OPERAND_BF16=1"] if bf16 else [])
                   + (["-DVG_PROJ_TRACE=1"] if trace else [])
                   + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(len(jobs)) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    if jobs or not os.path.exists(lib_path):
        run([_nvcc(), "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return lib_path


def build_all(force=False, verbose=False):
    return [build_library(force, verbose, bf16=False), build_library(force, verbose, bf16=True)]


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, trace=True))
    else:
        print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
