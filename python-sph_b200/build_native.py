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


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: tuning sweeps build VARIANTS (e.g. extra=["-DSPHB_MINB_NL_FORCE=6"]) into their own file, selected at
    run time with SPHB_LIB=<path>; the default library is only ever built with the default flags."""
    if out is None and not force and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    out = out or OUT
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    # python build_native.py [--force] [-v] [--out variant.so -DNAME=VALUE ...]
    args = sys.argv[1:]
    out = args[args.index("--out") + 1] if "--out" in args else None
    extra = [x for x in args if x.startswith("-D") or x.startswith("-maxrregcount")]
    print(build(force="--force" in args, verbose="-v" in args, out=out, extra=extra))
