"""ctypes wrapper of the TEST-ONLY host emulation (tests/emu/vm_emu.cpp) with the same
interface as robio_2021_dexopt_b200.engine, so parity checks are written once."""
import ctypes

import numpy as np


class EmuMemory:
    def __init__(self, lib, lanes):
        self.lib, self.lanes = lib, lanes
        self._h = ctypes.c_void_p()
        assert lib.emu_memory_create(lanes, ctypes.byref(self._h)) == 0, lib.emu_last_error()


class EmuExecutable:
    def __init__(self, lib, warps):
        self.lib, self.warps = lib, warps
        self._h = ctypes.c_void_p()
        self._compiled = False

    def compile(self, blob):
        buf = ctypes.create_string_buffer(blob, len(blob))
        st = self.lib.emu_program_create_from_blob(buf, len(blob), self.warps, ctypes.byref(self._h))
        if st != 0:
            raise RuntimeError("emu status %d: %s" % (st, self.lib.emu_last_error().decode()))
        self._compiled = True
        return self

    def _check(self):
        if not self._compiled:
            raise RuntimeError("call compile(...) before use")

    def check_streams(self):
        self._check()
        self.lib.emu_program_check_streams.argtypes = [ctypes.c_void_p]
        if self.lib.emu_program_check_streams(self._h) != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())

    def stat(self, key):
        v = ctypes.c_uint64()
        assert self.lib.emu_program_stat(self._h, key.encode(), ctypes.byref(v)) == 0
        return v.value

    def _bytes(self, which, lanes):
        v = ctypes.c_uint64()
        self.lib.emu_program_buffer_bytes(self._h, which, lanes, ctypes.byref(v))
        return v.value

    def inputBufferSize(self, lanes=8):
        return self._bytes(0, lanes)

    def outputBufferSize(self, lanes=8):
        return self._bytes(2, lanes)

    def parameterBufferSize(self, lanes=8):
        return self._bytes(1, lanes)

    def execute(self, memory):
        self._check()
        assert self.lib.emu_execute(self._h, memory._h) == 0

    def input(self, buffer, memory):
        self._check()
        b = np.ascontiguousarray(buffer, dtype=np.float64)
        st = self.lib.emu_input(self._h, memory._h, b.ctypes.data, b.nbytes)
        if st != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())

    def parameterize(self, buffer, memory):
        self._check()
        b = np.ascontiguousarray(buffer, dtype=np.float64)
        st = self.lib.emu_parameterize(self._h, memory._h, b.ctypes.data, b.nbytes)
        if st != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())

    def output(self, memory):
        self._check()
        out = np.empty(self._bytes(2, memory.lanes) // 8)
        st = self.lib.emu_output(self._h, memory._h, out.ctypes.data, out.nbytes)
        if st != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())
        return out

    inputVector = input
    parameterVector = parameterize

    def outputVector(self, memory):
        return self.output(memory)

    def run(self, v, memory):
        self.input(v, memory)
        self.execute(memory)
        return self.output(memory)


class EmuEngine:
    def __init__(self, path, warps=8):
        self.lib = ctypes.CDLL(path)
        c = ctypes
        self.lib.emu_last_error.restype = c.c_char_p
        self.lib.emu_program_create_from_blob.argtypes = [c.c_void_p, c.c_size_t, c.c_int, c.POINTER(c.c_void_p)]
        self.lib.emu_program_buffer_bytes.argtypes = [c.c_void_p, c.c_int, c.c_uint64, c.POINTER(c.c_uint64)]
        self.lib.emu_program_stat.argtypes = [c.c_void_p, c.c_char_p, c.POINTER(c.c_uint64)]
        self.lib.emu_memory_create.argtypes = [c.c_uint64, c.POINTER(c.c_void_p)]
        for n in ("emu_input", "emu_parameterize", "emu_output"):
            getattr(self.lib, n).argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64]
        self.lib.emu_execute.argtypes = [c.c_void_p, c.c_void_p]
        self.lib.emu_fused_create.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t,
                                              c.c_int, c.c_int, c.c_int, c.POINTER(c.c_void_p), c.POINTER(c.c_void_p)]
        self.warps = warps

    def direct_prepare(self, bundle):
        """(prep blob, bprop blob, stats) after the product's objective lowering step trx_direct.h"""
        c = ctypes
        f = self.lib.emu_direct_prepare
        f.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.POINTER(c.c_void_p), c.POINTER(c.c_size_t),
                      c.POINTER(c.c_void_p), c.POINTER(c.c_size_t), c.POINTER(c.c_uint64)]
        self.lib.emu_free.argtypes = [c.c_void_p]
        prep, bprop = bundle.programs["prep"], bundle.programs["bprop"]
        b0, b1 = c.create_string_buffer(prep, len(prep)), c.create_string_buffer(bprop, len(bprop))
        o0, o1, n0, n1 = c.c_void_p(), c.c_void_p(), c.c_size_t(), c.c_size_t()
        stats = (c.c_uint64 * 3)()
        if f(b0, len(prep), b1, len(bprop), c.byref(o0), c.byref(n0), c.byref(o1), c.byref(n1), stats) != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())
        out = (c.string_at(o0, n0.value), c.string_at(o1, n1.value),
               dict(prepare_dropped=stats[0], arguments_redirected=stats[1], mul_rewritten=stats[2]))
        self.lib.emu_free(o0)
        self.lib.emu_free(o1)
        return out

    def compile_fused(self, bundle, ring_slots=192, sync_penalty=900):
        """(forward+prepare, adjoint) executables from the objective's fused lowering"""
        blobs = [bundle.programs[n] for n in ("prog", "prep", "bprop")]
        bufs = [ctypes.create_string_buffer(b, len(b)) for b in blobs]
        fwd, adj = EmuExecutable(self.lib, self.warps), EmuExecutable(self.lib, self.warps)
        st = self.lib.emu_fused_create(bufs[0], len(blobs[0]), bufs[1], len(blobs[1]), bufs[2], len(blobs[2]),
                                       self.warps, ring_slots, sync_penalty, ctypes.byref(fwd._h), ctypes.byref(adj._h))
        if st != 0:
            raise RuntimeError("emu status %d: %s" % (st, self.lib.emu_last_error().decode()))
        fwd._compiled = adj._compiled = True
        return fwd, adj

    def createMemory(self, lanes=8):
        return EmuMemory(self.lib, lanes)

    def createExecutable(self):
        return EmuExecutable(self.lib, self.warps)

    def compile(self, blob):
        return self.createExecutable().compile(blob)
