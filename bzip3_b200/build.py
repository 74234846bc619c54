"""Builds the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbzip3_b200.so")
CLI = os.path.join(HERE, "bz3b200")            # file front end with the deep block queue (csrc/cli_main.cpp)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
         "-fPIC,-fvisibility=hidden", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp"))) + [
        os.path.join(os.path.dirname(HERE), "include", "libbz3.h"),
        os.path.join(os.path.dirname(HERE), "include", "bz3_b200.h")]


def stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "bz3_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libbzip3_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-o", CLI, os.path.join(CSRC, "cli_main.cpp"), "-ldl"])
    return LIB


def build_tools() -> None:
    """Extras for tuning sessions (git-ignored outputs under tools/variants/): the library with the clock64
    phase counters compiled in (-DBZ_CM_PROFILE, read by tools/cm_prof2.py) and the latency micro-benchmarks."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "tools", "variants")
    os.makedirs(out, exist_ok=True)
    for cmd in ([NVCC] + FLAGS + ["-DBZ_CM_PROFILE", "-o", os.path.join(out, "lib_cmprof.so"), os.path.join(CSRC, "bz3_api.cu")],
                [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", os.path.join(out, "ubench"),
                 os.path.join(root, "tools", "ubench.cu")],
                [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", os.path.join(out, "ubench_chain"),
                 os.path.join(root, "tools", "ubench_chain.cu")]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed: " + " ".join(cmd[-3:]))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--tools" in sys.argv:
        build_tools()
        print("built tools/variants/lib_cmprof.so, tools/variants/ubench and tools/variants/ubench_chain")
