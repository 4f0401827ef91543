"""Objective assembly: builds the synthetic dexopt rollout problem with the product's own tape
recorder and gradient derivation (C ABI trx_dexsyn_build, csrc/trx_builder.cpp).

Stands where dexopt does DexLearn::build + Solver::compile (dexopt/src/dexlearn.h:511-518,
tractor/src/core/solver.cpp:21-28)."""
from __future__ import annotations

import ctypes

from . import _native, trxfile


def build_dexsyn(**options) -> trxfile.Bundle:
    """options: kind, frames, outer, hidden, dropout, seed, joints, contacts, extra_fixed, wstd"""
    L = _native.lib()
    text = " ".join("%s=%s" % (k, v) for k, v in options.items())
    ptr = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_uint64()
    st = L.trx_dexsyn_build(text.encode(), ctypes.byref(ptr), ctypes.byref(n))
    if st != 0:
        raise _native.TrxError(st, L.trx_builder_last_error().decode(errors="replace"))
    try:
        data = ctypes.string_at(ptr, n.value)
    finally:
        L.trx_builder_free(ptr)
    return trxfile.loads(data)
