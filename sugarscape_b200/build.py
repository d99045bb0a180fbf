"""Builds the CUDA engine (sugarscape_b200/libssb.so) for sm_100a with nvcc, in-tree.

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "ssb_abi.cu")
OUT = os.environ.get("SSB_LIBSSB") or os.path.join(HERE, "libssb.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # fp64 parity with the reference's CPython arithmetic: no FMA contraction, IEEE div/sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(ROOT, "include", "ssb.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = []
    if os.environ.get("SSB_GROUP_LANES"):      # tuning builds: lanes cooperating on one agent (4, 8, 16, 32)
        extra.append("-DSSB_GROUP_LANES=" + os.environ["SSB_GROUP_LANES"])
    # mode E with decision models: the candidate cells are scored on all lanes (SSB_PARALLEL_ETHICS=0: on lane 0, the form the
    # host-compiled device source runs).  Round 2 on a B200: the 25 decision-model cases pass in both forms; determinism.json
    # through the drop-in driver runs 1.2-1.7 x faster in this one (profiles/r2_examples_driver_*.json)
    if os.environ.get("SSB_PARALLEL_ETHICS", "1") != "0":
        extra.append("-DSSB_PARALLEL_ETHICS")
    if os.environ.get("SSB_SLAB_BITS"):        # tuning builds: slabs of the sweep's priority order = 1 << bits
        extra.append("-DSSB_SLAB_BITS=" + os.environ["SSB_SLAB_BITS"])
    if os.environ.get("SSB_LEAN_BLOCKS"):      # tuning builds: resident blocks per SM the plain lean sweep kernel is compiled for
        extra.append("-DSSB_LEAN_BLOCKS=" + os.environ["SSB_LEAN_BLOCKS"])
    if os.environ.get("SSB_LEAN_PREFETCH"):
        extra.append("-DSSB_LEAN_PREFETCH=" + os.environ["SSB_LEAN_PREFETCH"])
    if os.environ.get("SSB_MARK_LANES"):
        extra.append("-DSSB_MARK_LANES=" + os.environ["SSB_MARK_LANES"])
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libssb.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(OUT)
