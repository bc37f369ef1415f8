#!/usr/bin/env python
"""Build libjabble_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python wobble_jax_b200/csrc/build.py [--force] [--verbose]

Output: wobble_jax_b200/libjabble_b200.so (git-ignored; travels to the GPU box with the snapshot).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libjabble_b200.so")
SOURCES = ["jb_plan.cu"]
DEPS = ["jb_plan.cu", "jb_kernels.cuh", "jb_fused2.cuh", "jb_fused32.cuh", "jb_strip.cuh", "jb_group.cuh", "jb_lbfgs.cuh", "jb_device.cuh", os.path.join(ROOT, "include", "jabble_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function,-fopenmp", "--shared", "-cudart", "shared", "-ldl", "-lgomp"]


def source_hash():
    h = hashlib.sha256()
    for d in DEPS:
        with open(os.path.join(HERE, d) if not os.path.isabs(d) else d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = OUT + ".hash"
    digest = source_hash()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    # built beside the target and moved into place: a reader (a snapshot of the tree, another process
    # loading the library) never sees a half-written file
    tmp = os.path.join(ROOT, "build", f"libjabble_b200.{os.getpid()}.so")
    os.makedirs(os.path.dirname(tmp), exist_ok=True)
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", tmp,
           *[os.path.join(HERE, s) for s in SOURCES]]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, OUT)
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
