"""In-tree build of libbrie_b200.so for sm_100a (`python -m brie_b200.build`)."""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "brie_abi.cu")
DEPS = [SRC, os.path.join(_HERE, "csrc", "brie_kernels.cuh"), os.path.join(_HERE, "csrc", "brie_philox.h"),
        os.path.join(os.path.dirname(_HERE), "include", "brie_b200.h")]
OUT = os.path.join(_HERE, "libbrie_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(p) <= t for p in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("BRIE_NVCC_EXTRA", "").split()
    out = os.environ.get("BRIE_LIB_OUT", OUT)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", out]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
