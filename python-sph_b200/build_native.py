"""Build libsphb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsphb200.so")
SOURCES = ["sphb_grid.cu", "sphb_gather.cu", "sphb_force.cu", "sphb_nlist.cu"]
HEADERS = ["sphb_common.cuh", "sphb_scan.cuh", "sphb_pair.cuh", os.path.join("..", "..", "include", "sphb200.h")]
# -fmad=false: only the fma() calls written in the sources are fused.  Implicit contraction would let two inlined
# copies of the same pair function (list drained inside the scan / after it, even / odd unrolled slot) round
# differently, and per-particle sums would then depend on how targets are chunked into warps and ranks.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("SPHB_NVCC_EXTRA", "").split()   # tuning sweeps only (e.g. -DSPHB_LIST=96)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
