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


TRANSLATION_UNITS = ["trx_engine.cu", "trx_rollout.cu", "trx_builder.cpp"]
OBJ_DIR = os.path.join(HERE, "_obj")


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


def _unit_deps(unit: str):
    """files a translation unit is rebuilt for: found by scanning its #include "..." closure"""
    import re
    seen, todo = set(), [os.path.join(CSRC, unit)]
    while todo:
        path = os.path.normpath(todo.pop())
        if path in seen or not os.path.exists(path):
            continue
        seen.add(path)
        with open(path, errors="replace") as f:
            for inc in re.findall(r'#\s*include\s*"([^"]+)"', f.read()):
                todo.append(os.path.join(os.path.dirname(path), inc))
    return seen


def build_native(force: bool = False, verbose: bool = False) -> str:
    """one object per translation unit (compiled in parallel, rebuilt only when its include closure
    changed), then one link: csrc/*.cu|cpp -> _obj/*.o -> libtrx_b200.so"""
    if not force and not stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-c"]
    # development builds: TRX_NVCC_EXTRA="-DTRX_ROLLOUT_PROFILE" adds the per-phase clocks of csrc/trx_rollout.cu
    compile_flags += os.environ.get("TRX_NVCC_EXTRA", "").split()
    if verbose:
        compile_flags = ["-Xptxas", "-v"] + compile_flags
    jobs, objs = [], []
    for unit in TRANSLATION_UNITS:
        obj = os.path.join(OBJ_DIR, os.path.splitext(unit)[0] + ".o")
        objs.append(obj)
        deps = _unit_deps(unit)
        fresh = os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps)
        if fresh and not force:
            continue
        jobs.append((unit, subprocess.Popen([nvcc] + compile_flags + [os.path.join(CSRC, unit), "-o", obj], cwd=CSRC)))
    for unit, job in jobs:
        if job.wait() != 0:
            raise subprocess.CalledProcessError(job.returncode, "nvcc " + unit)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", OUT + ".tmp"])
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
