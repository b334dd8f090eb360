"""Builds csrc/libresenc_b200.so (the C ABI of include/resenc_b200.h) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libresenc_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "conv_tc5.cuh", "conv_tc5t.cuh", "conv_slab.cuh", "loss.cuh", "conv_generic.cuh", "wgrad_tc5.cuh", "wgrad2_tc5.cuh", "elementwise.cuh", "split.cuh", "blend.cuh", "optim.cuh",
           os.path.join("..", "..", "include", "resenc_b200.h")]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-cudart", "static"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, f)) <= t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [_nvcc(), *FLAGS, "-o", OUT, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
