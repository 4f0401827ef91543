"""Parity checks shared by the CPU (emulated lowering + op bodies) and GPU (CUDA engine) test files.

Every expected value comes from the REFERENCE's own SimpleEngine via the golden fixtures
(tests/golden/generate.sh).  Tolerance: the north star asks for 1e-9 relative in fp64; the
checks below use 1e-11 relative (+1e-13 absolute for values that are exact zeros in the
reference) so that FMA contraction on the GPU is the only slack.
"""
import numpy as np

from robio_2021_dexopt_b200 import trxfile

RTOL = 1e-11
ATOL = 1e-13


def assert_close(got, want, what, rtol=RTOL, atol=ATOL):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    scale = max(1.0, float(np.max(np.abs(want))) if want.size else 1.0)
    err = np.abs(got - want)
    bad = err > atol * scale + rtol * np.abs(want)
    if np.any(bad):
        i = int(np.argmax(err))
        raise AssertionError("%s: %d/%d elements differ, worst at %d: got %.17g want %.17g (err %.3g)" %
                             (what, int(bad.sum()), bad.size, i, got.flat[i], want.flat[i], err.flat[i]))


def op_names(bundle):
    return [line.split() for line in bundle.texts["ops"].splitlines() if line.strip()]


def check_single_op(engine, bundle, name, has_grad):
    """tests the way tractor/src/test/test.cpp:565-616 drives one operator"""
    pre = "op/%s/" % name
    x_prog = engine.compile(bundle.programs[pre + "prog"])
    mem = engine.createMemory(8)
    y = x_prog.run(bundle.arrays[pre + "x"], mem)
    assert_close(y, bundle.arrays[pre + "y"], name + " value")
    if not has_grad:
        return
    x_prep = engine.compile(bundle.programs[pre + "prep"])
    x_fprop = engine.compile(bundle.programs[pre + "fprop"])
    x_bprop = engine.compile(bundle.programs[pre + "bprop"])
    x_prep.execute(mem)
    gx = x_bprop.run(bundle.arrays[pre + "gy"], mem)
    assert_close(gx, bundle.arrays[pre + "gx"], name + " reverse")
    mem = engine.createMemory(8)
    x_prog.run(bundle.arrays[pre + "x"], mem)
    x_prep.execute(mem)
    dy = x_fprop.run(bundle.arrays[pre + "dx"], mem)
    assert_close(dy, bundle.arrays[pre + "dy"], name + " forward")


def tile_arrays(bundle, key):
    out = []
    t = 0
    while "tile%d/%s" % (t, key) in bundle.arrays:
        out.append(bundle.arrays["tile%d/%s" % (t, key)])
        t += 1
    return out


def check_problem(engine, bundle, n_tiles=None, check_tangent=True, compile_bprop=None, blobs=None):
    """objective + gradient over several lane tiles: x_prog, x_prep, x_bprop as in solvers/gd.h:50-63.
    compile_bprop: alternative way to lower the adjoint tape (e.g. with accumulate folding)"""
    prog = bundle.view("prog")
    params = tile_arrays(bundle, "params")
    r_ref = tile_arrays(bundle, "r")
    if n_tiles is None:
        n_tiles = len(params)
    params, r_ref = params[:n_tiles], r_ref[:n_tiles]
    lanes = 8 * n_tiles
    blobs = dict(bundle.programs, **(blobs or {}))  # blobs: rewritten tapes replacing the bundle's
    x_prog = engine.compile(blobs["prog"])
    x_prep = engine.compile(blobs["prep"])
    x_bprop = (compile_bprop or engine.compile)(blobs["bprop"])
    mem = engine.createMemory(lanes)
    if prog.parameters:
        x_prog.parameterize(trxfile.tiles_to_wide(prog.parameters, params), mem)
    x = bundle.arrays["x"]
    r = x_prog.run(x, mem)
    r_want = trxfile.tiles_to_wide(prog.outputs, r_ref)
    assert_close(r, r_want, "residuals")
    loss = float(np.dot(r, r))
    if n_tiles == len(tile_arrays(bundle, "r")):
        want = float(bundle.arrays["loss_total"][0])
        assert abs(loss - want) <= 1e-11 * abs(want), "loss %.17g vs %.17g" % (loss, want)
    x_prep.execute(mem)
    g = x_bprop.run(r, mem)  # g = J^T r ; wide r has the scalar-typed rows once
    if n_tiles == len(tile_arrays(bundle, "r")):
        assert_close(g, bundle.arrays["g_total"], "gradient")
    elif n_tiles == 1:
        assert_close(g, bundle.arrays["tile0/g"], "gradient (tile 0)")
    if check_tangent:
        fprop = bundle.view("fprop")
        x_fprop = engine.compile(bundle.programs["fprop"])
        dr = x_fprop.run(bundle.arrays["dx"], mem)
        dr_want = trxfile.tiles_to_wide(fprop.outputs, tile_arrays(bundle, "dr")[:n_tiles])
        assert_close(dr, dr_want, "tangent")
    return r, g


def check_gd_iterates(make_step, bundle, steps=10):
    """first `steps` GradientDescentSolver iterates against the reference's (solvers/gd.h:43-129,
    lr 0.01, momentum 0 as dexopt.cpp:199-202); tolerance 1e-9 relative as the north star asks"""
    nx = bundle.arrays["x"].size
    want_x = bundle.arrays["gd/x"].reshape(-1, nx)
    want_loss = bundle.arrays["gd/loss"]
    assert want_x.shape[0] >= steps
    for k in range(steps):
        loss, x = make_step()
        assert abs(loss - want_loss[k]) <= 1e-9 * abs(want_loss[k]), "GD step %d loss %.17g vs %.17g" % (k, loss, want_loss[k])
        assert_close(x, want_x[k], "GD iterate %d" % (k + 1), rtol=1e-9, atol=1e-12)


class ExecutableObjective:
    """objective + gradient written against the Executable interface only (works on any engine)"""

    def __init__(self, engine, bundle, n_tiles):
        self.prog = bundle.view("prog")
        self.x_prog = engine.compile(bundle.programs["prog"])
        self.x_prep = engine.compile(bundle.programs["prep"])
        self.x_bprop = engine.compile(bundle.programs["bprop"])
        self.mem = engine.createMemory(8 * n_tiles)
        params = tile_arrays(bundle, "params")[:n_tiles]
        if self.prog.parameters:
            self.x_prog.parameterize(trxfile.tiles_to_wide(self.prog.parameters, params), self.mem)

    def evaluate(self, x):
        r = self.x_prog.run(x, self.mem)
        self.x_prep.execute(self.mem)
        g = self.x_bprop.run(r, self.mem)
        return float(np.dot(r, r)), g
