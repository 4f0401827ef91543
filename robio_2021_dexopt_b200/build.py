"""In-tree build of the native CUDA engine (sm_100a only).

    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... csrc/trx_engine.cu -> libtrx_b200.so

The .so stays in the package directory so that it travels with the repo snapshot and is
visible as an in-tree native library to whoever records loaded objects.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtrx_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # the interpreter's opcode switch is dense: let NVVM emit one brx.idx jump table for it instead of a
    # compare tree in front of small tables (the default density threshold of 101 % never does)
    "--jump-table-density=25",
]


def sources():
    deps = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".def", ".cpp")):
                deps.append(os.path.join(root, f))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "trx_b200.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "trx_file.h"))
    return deps


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, "trx_engine.cu"), os.path.join(CSRC, "trx_builder.cpp"),
                                 "-o", OUT + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd, cwd=CSRC)
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
