"""Build ``pysdc_b200/lib/libsdcb200.so`` with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pysdc_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libsdcb200.so")
SOURCES = ["colloc.cu", "stencil.cu", "cg.cu", "highorder.cu", "gmres.cu", "reaction.cu", "direct.cu", "peer.cu", "transfer.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]


def _newest_source():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sdc_b200.h")]
    return max(os.path.getmtime(f) for f in files if os.path.isfile(f))


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    extra = os.environ.get("SDCB200_NVCC_DEFS", "").split()  # e.g. -DSDCB200_TWO_CTAS (A/B experiments)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
