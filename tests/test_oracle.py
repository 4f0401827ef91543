"""Pins the numpy restatement (oracle/tape_oracle.py) to the golden vectors produced by the reference's
own SimpleEngine (oracle/_ref/ref_tool, see tests/golden/generate.sh)."""
import os
import sys

import numpy as np
import pytest

import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tape_oracle as oracle  # noqa: E402


class OracleEngine:
    """adapts the oracle to the engine interface the shared parity checks use (one lane tile)"""

    class X:
        def __init__(self, view):
            self.x = oracle.OracleExecutable(view)

        def run(self, v, mem):
            return self.x.run(v, mem)

        def execute(self, mem):
            self.x.execute(mem)

        def parameterize(self, v, mem):
            self.x.parameterize(v, mem)

    def compile(self, blob):
        from robio_2021_dexopt_b200 import trxfile
        return OracleEngine.X(trxfile.parse_program(blob))

    def createMemory(self, lanes=8):
        assert lanes == 8
        return oracle.OracleMemory()


@pytest.mark.parametrize("fixture", ["optests_8d.trxb", "optests_d.trxb"])
def test_every_operator(golden, fixture):
    bundle = golden(fixture)
    eng = OracleEngine()
    for name, grad in parity.op_names(bundle):
        parity.check_single_op(eng, bundle, name, grad == "grad")


@pytest.mark.parametrize("fixture", ["fkchain24.trxb", "dexsyn_small.trxb.xz"])
def test_problem_tile0(golden, fixture):
    parity.check_problem(OracleEngine(), golden(fixture), n_tiles=1)


def test_two_tiles_and_gradient_descent(golden):
    """per-tile residuals, summed gradient with the lane-independent residual rows counted once, and the
    first GD iterates (solvers/gd.h:43-129)"""
    bundle = golden("fkchain24.trxb")
    prog = bundle.view("prog")
    xs = {n: oracle.OracleExecutable(bundle.view(n)) for n in ("prog", "prep", "bprop")}
    params = parity.tile_arrays(bundle, "params")
    scalar = np.concatenate([np.full(p.size // 8, p.lanes == 1) for p in prog.outputs])
    lr = float(bundle.arrays["gd/learning_rate"][0])
    x = bundle.arrays["x"].copy()
    nx = x.size
    want_x = bundle.arrays["gd/x"].reshape(-1, nx)
    for k in range(4):
        g_sum = np.zeros(nx)
        loss = 0.0
        for t, p in enumerate(params):
            mem = oracle.OracleMemory()
            xs["prog"].parameterize(p, mem)
            r = xs["prog"].run(x, mem)
            if k == 0:
                parity.assert_close(r, bundle.arrays["tile%d/r" % t], "tile %d residuals" % t)
            if t > 0:
                r = np.where(scalar, 0.0, r)
            loss += float(r @ r)
            xs["prep"].execute(mem)
            g_sum += xs["bprop"].run(r, mem)
        if k == 0:
            parity.assert_close(g_sum, bundle.arrays["g_total"], "gradient")
        assert abs(loss - bundle.arrays["gd/loss"][k]) <= 1e-11 * abs(loss)
        x = x + (-lr * g_sum)
        parity.assert_close(x, want_x[k], "GD iterate %d" % (k + 1), rtol=1e-10)
