"""Builds libmunk_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library has no torch dependency; Python reaches it through ctypes (munk_b200/_lib.py).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libmunk_b200.so")
SOURCES = ["api.cu", "dgemm.cu", "factor.cu", "leaf.cu", "diag.cu", "trsm_strip.cu", "assemble.cu", "scores.cu", "jacobi.cu",
           "dist.cu", "hostio.cu", "microbench.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "--fmad=true",
]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into lib/libmunk_b200.so; returns the path."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "include", "munk_b200.h"))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest(deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # dist.cu includes <nccl.h> for the types only; NCCL itself is dlopen'ed at run time (-ldl)
    cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-o", OUT] + srcs + ["-ldl", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmunk_b200.so")
    if verbose:
        print(res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
