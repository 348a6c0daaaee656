"""Builds libe1b200.so (the C-ABI + sm_100a kernels) in-tree with nvcc.

    python galileo-sdr-sim_b200/build.py [--force]

The library lands in galileo-sdr-sim_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
nvcc cross-compiles for sm_100a, so this works on a machine without a GPU.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "e1b200_capi.cu"
DEPS = [SRC, PKG / "csrc" / "e1_kernels.cuh", PKG / "csrc" / "e1_core.h", PKG / "data" / "e1_prn_codes.h",
        PKG.parent / "include" / "e1b200.h"]
LIB = PKG / "lib" / "libe1b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the exact paths use explicit __dadd_rn/__dmul_rn; switch contraction off everywhere anyway
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def build_lib(force=False, verbose=False, defines=(), out=None):
    """defines / out: tuning builds (extra -D flags, alternative output path; select one at run time with E1B200_LIB=<path>)."""
    lib = Path(out) if out else LIB
    if not force and lib.exists() and all(lib.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return lib
    lib.parent.mkdir(exist_ok=True)
    cmd = ([nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines]
           + ["-o", str(lib), str(SRC)])
    subprocess.check_call(cmd)
    return lib


HOST_SRC = PKG / "host" / "e1_scenario.cpp"
FIFO_SRC = PKG / "host" / "e1_fifo.cpp"
HOST_LIB = PKG / "lib" / "libe1host.so"
CLI_SRC = PKG / "host" / "e1sim_main.cpp"
CLI_BIN = PKG / "lib" / "e1sim"


def build_host(force=False):
    """Host side of the drop-in (RINEX -> records, plain C++, no CUDA): lib/libe1host.so.
    -ffp-contract=off: the records must be bit-identical to the reference's (see e1_scenario.h)."""
    deps = [HOST_SRC, FIFO_SRC, PKG / "host" / "e1_scenario.h", PKG / "host" / "e1_fifo.h", PKG / "csrc" / "e1_core.h",
            PKG.parent / "include" / "e1b200.h"]
    if not force and HOST_LIB.exists() and all(HOST_LIB.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return HOST_LIB
    HOST_LIB.parent.mkdir(exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", str(HOST_LIB), str(HOST_SRC), str(FIFO_SRC),
                           "-lpthread"])
    return HOST_LIB


def build_cli(force=False):
    """lib/e1sim: the reference's command line (src/main.cpp:216) over libe1host + libe1b200."""
    build_lib()
    build_host()
    deps = [CLI_SRC, HOST_LIB, LIB]
    if not force and CLI_BIN.exists() and all(CLI_BIN.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return CLI_BIN
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", str(CLI_BIN), str(CLI_SRC), "-L", str(LIB.parent), "-le1host", "-le1b200",
                           "-lpthread", "-Wl,-rpath,$ORIGIN"])
    return CLI_BIN


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
