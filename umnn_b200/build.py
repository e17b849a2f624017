"""Builds umnn_b200/libumnn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m umnn_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libumnn_b200.so")
SOURCES = ["umnn_abi.cu", "cc_forward_fp32.cu", "cc_forward_tc.cu", "cc_backward_fp32.cu", "cc_backward_tc.cu", "invert_bracket.cu"]
HEADERS = ["umnn_common.cuh", "tc_common.cuh", "tc_layout.cuh", "tc_bwd_layout.cuh", "tc_kernels.cuh", os.path.join("..", "..", "include", "umnn_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = LIB_PATH) -> str:
    """Compile every CUDA source into one shared library; returns its path.  `extra_flags` / `out` build an
    experiment variant next to the product library (see scripts/build_variants.py)."""
    if not force and not extra_flags and out == LIB_PATH and not _stale():
        return LIB_PATH
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if os.environ.get("UMNN_B200_TC_SPIN_LIMIT"):
        # override the bound on mbarrier waits in nanoseconds (default 2^32 ~ 4.3 s, see tc_common.cuh); 0 = wait forever
        flags += ["-DUMNN_TC_SPIN_LIMIT=" + os.environ["UMNN_B200_TC_SPIN_LIMIT"] + "LL"]
    cmd = [_nvcc()] + flags + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libumnn_b200.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
