"""Host-side mirror of tractor's Engine / Executable / Memory interface on the B200 engine.

Reference interface being mirrored (same method names, argument meaning and error behaviour):
    tractor/include/tractor/core/engine.h:11-112
        struct Memory;  class Executable { compile, execute, input, output, parameterize,
        inputVector, outputVector, run, parameterVector, inputBufferSize, outputBufferSize };
        struct Engine { createMemory, createExecutable, compile }
    tractor/src/core/engine.cpp:9-78 (wrappers, "call compile(...) before use")

Differences that come from running R rollouts instead of 8 SIMD lanes:
    * Memory is created for a number of lanes R (multiple of 8); buffers use the
      wide packed layout documented in include/trx_b200.h (identical to the
      reference's Buffer for R == 8).
    * A program is handed over in its serialised neutral form (TRXB program
      payload, include/trx_file.h) instead of as a C++ tractor::Program object;
      INTEGRATION.md shows the C++ shim that fills the same C ABI from a
      tractor::Program directly.

Everything here is plumbing over the C ABI (include/trx_b200.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _native
from ._native import TrxError, check

BUF_INPUT, BUF_PARAMETER, BUF_OUTPUT = 0, 1, 2


class Memory:
    """tractor::Memory (engine.h:11-13): opaque, engine-defined state shared by a solver's executables."""

    def __init__(self, lanes: int, device: int = 0):
        self._h = ctypes.c_void_p()
        self.lanes = int(lanes)
        self.device = device
        check(_native.lib().trx_memory_create(self.lanes, device, ctypes.byref(self._h)))

    def bytes(self) -> int:
        v = ctypes.c_uint64()
        check(_native.lib().trx_memory_bytes(self._h, ctypes.byref(v)))
        return v.value

    def peek(self, tile: int, address: int, n_doubles: int) -> np.ndarray:
        out = np.empty(n_doubles)
        check(_native.lib().trx_memory_peek(self._h, tile, address, out.ctypes.data, out.nbytes))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _native.lib().trx_memory_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:  # interpreter teardown
            pass


class Executable:
    """tractor::Executable (engine.h:15-106)."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        self._compiled = False
        self.device = device

    def _check_compiled(self):
        if not self._compiled:
            raise RuntimeError("call compile(...) before use")  # engine.h:20-24

    # -- compile -------------------------------------------------------------
    def compile(self, program_blob: bytes) -> "Executable":
        if self._h.value:
            _native.lib().trx_program_destroy(self._h)
            self._h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(program_blob, len(program_blob))
        check(_native.lib().trx_program_create_from_blob(buf, len(program_blob), self.device, ctypes.byref(self._h)))
        self._compiled = True
        return self

    def stat(self, key: str) -> int:
        self._check_compiled()
        v = ctypes.c_uint64()
        check(_native.lib().trx_program_stat(self._h, key.encode(), ctypes.byref(v)))
        return v.value

    def _bytes(self, which: int, lanes: int) -> int:
        v = ctypes.c_uint64()
        check(_native.lib().trx_program_buffer_bytes(self._h, which, lanes, ctypes.byref(v)))
        return v.value

    def inputBufferSize(self, lanes: int = 8) -> int:
        self._check_compiled()
        return self._bytes(BUF_INPUT, lanes)

    def outputBufferSize(self, lanes: int = 8) -> int:
        self._check_compiled()
        return self._bytes(BUF_OUTPUT, lanes)

    def parameterBufferSize(self, lanes: int = 8) -> int:
        self._check_compiled()
        return self._bytes(BUF_PARAMETER, lanes)

    # -- the five virtuals of the reference ------------------------------------
    def execute(self, memory: Memory, stream: int = 0) -> None:
        self._check_compiled()
        check(_native.lib().trx_execute(self._h, memory._h, ctypes.c_void_p(stream)))

    def input(self, buffer: np.ndarray, memory: Memory, stream: int = 0) -> None:
        self._check_compiled()
        b = np.ascontiguousarray(buffer, dtype=np.float64)
        check(_native.lib().trx_input(self._h, memory._h, b.ctypes.data, b.nbytes, ctypes.c_void_p(stream)))

    def parameterize(self, buffer: np.ndarray, memory: Memory, stream: int = 0) -> None:
        self._check_compiled()
        b = np.ascontiguousarray(buffer, dtype=np.float64)
        check(_native.lib().trx_parameterize(self._h, memory._h, b.ctypes.data, b.nbytes, ctypes.c_void_p(stream)))

    def output(self, memory: Memory, out: Optional[np.ndarray] = None, stream: int = 0) -> np.ndarray:
        self._check_compiled()
        n = self._bytes(BUF_OUTPUT, memory.lanes) // 8
        if out is None:
            out = np.empty(n)
        assert out.dtype == np.float64 and out.size == n and out.flags.c_contiguous
        check(_native.lib().trx_output(self._h, memory._h, out.ctypes.data, out.nbytes, ctypes.c_void_p(stream)))
        return out

    # -- device-pointer variants (wide layout, asynchronous on `stream`) ----------
    def input_device(self, ptr: int, nbytes: int, memory: Memory, stream: int = 0) -> None:
        self._check_compiled()
        check(_native.lib().trx_input_device(self._h, memory._h, ctypes.c_void_p(ptr), nbytes, ctypes.c_void_p(stream)))

    def parameterize_device(self, ptr: int, nbytes: int, memory: Memory, stream: int = 0) -> None:
        self._check_compiled()
        check(_native.lib().trx_parameterize_device(self._h, memory._h, ctypes.c_void_p(ptr), nbytes,
                                                    ctypes.c_void_p(stream)))

    def output_device(self, ptr: int, nbytes: int, memory: Memory, stream: int = 0) -> None:
        self._check_compiled()
        check(_native.lib().trx_output_device(self._h, memory._h, ctypes.c_void_p(ptr), nbytes, ctypes.c_void_p(stream)))

    # -- helpers with the reference's names (engine.h:69-105) ---------------------
    def inputVector(self, v, memory: Memory) -> None:
        self.input(np.asarray(v, dtype=np.float64), memory)

    def outputVector(self, memory: Memory) -> np.ndarray:
        return self.output(memory)

    def parameterVector(self, v, memory: Memory) -> None:
        self.parameterize(np.asarray(v, dtype=np.float64), memory)

    def run(self, v, memory: Memory) -> np.ndarray:
        self.inputVector(v, memory)
        self.execute(memory)
        return self.outputVector(memory)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _native.lib().trx_program_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:  # interpreter teardown
            pass


class CudaEngine:
    """tractor::Engine (engine.h:108-112) on one B200."""

    def __init__(self, device: int = 0):
        self.device = device
        if _native.lib().trx_device_count() <= 0:
            raise TrxError(-3, "no CUDA device: the tape engine has no CPU execution path")

    def createMemory(self, lanes: int = 8) -> Memory:
        return Memory(lanes, self.device)

    def createExecutable(self) -> Executable:
        return Executable(self.device)

    def compile(self, program_blob: bytes) -> Executable:
        x = self.createExecutable()
        x.compile(program_blob)
        return x


def kernel_launch_count() -> int:
    return int(_native.lib().trx_kernel_launch_count())
