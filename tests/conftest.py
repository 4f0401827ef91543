import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_emulator():
    """TEST-ONLY host emulation of the VM (tests/emu/vm_emu.cpp); never loaded by the package."""
    out = os.path.join(ROOT, "tests", "_build", "libvm_emu.so")
    srcs = [
        os.path.join(ROOT, "tests", "emu", "vm_emu.cpp"),
        os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc", "vm_ops.def"),
        os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc", "vm_exec.cuh"),
        os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc", "vm_math.cuh"),
        os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc", "collision", "trx_collide.h"),
        os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc", "collision", "trx_shape_table.h"),
        os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc", "trx_lower.h"),
        os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc", "trx_ops.h"),
        os.path.join(ROOT, "include", "trx_file.h"),
    ]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", srcs[0], "-o", out])
    return out


@pytest.fixture(scope="session")
def emu_engine():
    from emu_engine import EmuEngine
    return EmuEngine(_build_emulator())


@pytest.fixture(scope="session")
def cuda_engine():
    import robio_2021_dexopt_b200 as pkg
    return pkg.CudaEngine(0)


@pytest.fixture(scope="session")
def golden():
    from robio_2021_dexopt_b200 import trxfile
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = trxfile.load(os.path.join(GOLDEN, name))
        return cache[name]

    return get
