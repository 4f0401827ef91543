"""Helpers shared by the rollout-objective tests: golden fixtures -> wide layouts, and the ctypes wrapper of the
TEST-ONLY host execution of the frame phases (tests/emu/rollout_emu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# fixture -> problem options (tests/golden/generate.sh)
FIXTURES = {
    "rollout_handarm_h64.trxb": dict(kind="handarm", hidden=64, frames=4),
    "rollout_hand8_linear.trxb": dict(kind="hand8", hidden=0, frames=5),
    "rollout_chain30.trxb": dict(kind="chain", joints=30, contacts=16, hidden=16, frames=3),
    "rollout_handarm_long.trxb": dict(kind="handarm", hidden=8, frames=24),
}


def option_text(options):
    return " ".join("%s=%s" % (k, v) for k, v in options.items())


def wide_from_tiles(bundle, n_scalar):
    """golden tiles ([ports][components][8 lanes] per tile, scalar rows first) -> wide layout:
    params [3][lanes], r = scalar rows once, then [row][lanes]"""
    params = parity.tile_arrays(bundle, "params")
    rt = parity.tile_arrays(bundle, "r")
    pw = np.concatenate([p.reshape(-1, 8) for p in params], axis=1).copy()
    rows = (rt[0].size - n_scalar) // 8
    lane_part = np.concatenate([r[n_scalar:].reshape(rows, 8) for r in rt], axis=1)
    r = np.concatenate([rt[0][:n_scalar], lane_part.ravel()])
    return pw, r


def build_rollout_emulator():
    out = os.path.join(ROOT, "tests", "_build", "librollout_emu.so")
    csrc = os.path.join(ROOT, "robio_2021_dexopt_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "emu", "rollout_emu.cpp"),
            os.path.join(csrc, "rollout", "rollout_frame.h"), os.path.join(csrc, "rollout", "rollout_plan.h"),
            os.path.join(csrc, "assembly", "dexsyn_model.h"), os.path.join(csrc, "assembly", "dexsyn_options.h"),
            os.path.join(csrc, "vm_math.cuh")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", srcs[0], "-o", out])
    return out


class RolloutEmu:
    def __init__(self):
        self.lib = ctypes.CDLL(build_rollout_emulator())
        self.lib.rollout_emu_last_error.restype = ctypes.c_char_p

    def info(self, options, key):
        v = ctypes.c_long()
        st = self.lib.rollout_emu_info(option_text(options).encode(), key.encode(), ctypes.byref(v))
        assert st == 0, self.lib.rollout_emu_last_error()
        return v.value

    def evaluate(self, options, params_wide, x, want_r=True, want_g=True, seed=None):
        lanes = params_wide.shape[1]
        ns, rows, nx = (self.info(options, k) for k in ("n_scalar", "lane_rows", "n_x"))
        dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
        pw = np.ascontiguousarray(params_wide, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(ns + rows * lanes) if want_r else None
        g = np.zeros(nx) if want_g else None
        loss = ctypes.c_double()
        st = self.lib.rollout_emu_eval(option_text(options).encode(), lanes, dp(pw), dp(x),
                                       dp(r) if want_r else None, ctypes.byref(loss), dp(g) if want_g else None,
                                       dp(np.ascontiguousarray(seed)) if seed is not None else None)
        assert st == 0, self.lib.rollout_emu_last_error()
        return r, loss.value, g
