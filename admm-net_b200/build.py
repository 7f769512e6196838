"""Build libadmmnet_b200.so in-tree with nvcc for sm_100a (no torch involved; plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libadmmnet_b200.so")
SRC = [os.path.join(HERE, "csrc", f) for f in ("capi.cu", "net_kernels.cu", "arrow_kernels.cu", "classic_kernels.cu", "peak_kernels.cu", "gen_kernels.cu", "head_kernels.cu", "tc_probe.cu", "tc.cuh", "tail_tc.cu", "big_kernels.cu", "dc_kernels.cu", "trd_reg.cuh",
                                                "common.cuh")]
HDR = os.path.join(os.path.dirname(HERE), "include", "admmnet_b200.h")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SRC + [HDR])


def build(force=False, verbose=False, debug=False):
    """debug=True adds -DADMMNET_DEBUG: in-kernel bounds asserts (csrc/common.cuh ADMM_ASSERT) that trap with file:line."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", os.path.join(HERE, "csrc", "capi.cu"), "-o", LIB]
    if debug:
        cmd.insert(1, "-DADMMNET_DEBUG")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv or "--debug" in sys.argv, verbose=True, debug="--debug" in sys.argv)
