"""Builds libb200aabb.so (CUDA, sm_100a) in-tree with nvcc.  `python -m cgal_b200._build` or __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200aabb.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-warn-spills"]


def sources():
    inc = os.path.join(os.path.dirname(HERE), "include")
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(inc, "b200aabb.h"), os.path.join(inc, "b200work.h")]


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    extra = os.environ.get("B200AABB_NVCC_EXTRA", "").split()      # tuning experiments only (scripts/), e.g. -DB2_PP_WARPS_PER_SM=16
    cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-o", LIB, os.path.join(CSRC, "b200aabb.cu"), os.path.join(CSRC, "workload.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
