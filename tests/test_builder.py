"""The product's own tape recorder + gradient derivation (csrc/builder) against the reference's
recording of the same assembly (tests/golden, recorded through the reference Var<> API and
differentiated by the reference buildGradients)."""
import numpy as np
import pytest

import parity
from robio_2021_dexopt_b200 import trxfile

CASES = [
    ("dexsyn_small.trxb.xz", dict(frames=3, hidden=4, extra_fixed=4)),
    ("dexsyn_dropout_outer2.trxb.xz", dict(frames=2, hidden=4, extra_fixed=2, dropout=1, outer=2)),
]


def _histogram(view):
    h = {}
    nargs = [len(o.args) for o in view.ops]
    pc = 0
    code = view.code
    while pc < len(code):
        oi = int(code[pc])
        h[view.ops[oi].name] = h.get(view.ops[oi].name, 0) + 1
        pc += 1 + nargs[oi]
    return h


@pytest.mark.parametrize("fixture,options", CASES)
def test_recorded_tape_equals_reference_recording(golden, fixture, options):
    """same operators, same counts; the reference additionally records one move per goal
    (Recorder::outputImpl, core/recorder.h:31-45) which the product's recorder does not need"""
    import robio_2021_dexopt_b200 as pkg
    ref = golden(fixture)
    mine = pkg.build_dexsyn(**options)
    np.testing.assert_array_equal(mine.arrays["x"], ref.arrays["x"])
    hm, hr = _histogram(mine.view("prog")), _histogram(ref.view("prog"))
    n_goal_moves = 0
    for name in sorted(set(hm) | set(hr)):
        if name.startswith("move"):
            n_goal_moves += hr.get(name, 0) - hm.get(name, 0)
        else:
            assert hm.get(name, 0) == hr.get(name, 0), name
    assert n_goal_moves == len(ref.view("prog").outputs)
    assert len(mine.view("prog").outputs) == len(ref.view("prog").outputs)
    assert len(mine.view("prog").inputs) == len(ref.view("prog").inputs)
    assert len(mine.view("prog").parameters) == len(ref.view("prog").parameters)


@pytest.mark.parametrize("fixture,options", CASES)
@pytest.mark.parametrize("fuse_dense", [0, 1])
def test_built_programs_reproduce_reference_numbers(emu_engine, golden, fixture, options, fuse_dense):
    """residuals, loss, gradient (and tangent) of the product-built programs equal the reference engine's
    results on the reference-recorded tape; fuse_dense=1 records dense layers as block instructions"""
    import robio_2021_dexopt_b200 as pkg
    ref = golden(fixture)
    mine = pkg.build_dexsyn(fuse_dense=fuse_dense, **options)
    for k, v in ref.arrays.items():
        mine.arrays.setdefault(k, v)
    parity.check_problem(emu_engine, mine, check_tangent=True)  # fused tapes: forward_x_dense block operator
    if fuse_dense:
        n_ref = int(ref.meta()["n_prog_instructions"])
        assert int(mine.meta()["n_prog_instructions"]) < 0.7 * n_ref
        assert any(o.name == "forward_x_dense_8d" for o in mine.view("fprop").ops)
        assert any(o.name == "forward_x_dense_8d" for o in mine.view("hprop").ops)


@pytest.mark.parametrize("fuse_dense", [0, 1])
def test_hprop_is_bprop_after_fprop(emu_engine, golden, fuse_dense):
    """x_hprop = x_fprop ; x_bprop on one memory (gradients.cpp:344-362): J^T J v from the single program equals
    the two sweeps run one after the other, for the reference expansion and for dense-fused tapes
    (forward_x_dense / reverse_x_dense)"""
    import numpy as np
    import robio_2021_dexopt_b200 as pkg
    from robio_2021_dexopt_b200 import trxfile
    ref = golden("dexsyn_small.trxb.xz")
    mine = pkg.build_dexsyn(frames=3, hidden=4, extra_fixed=4, fuse_dense=fuse_dense)
    x = {n: emu_engine.compile(mine.programs[n]) for n in ("prog", "prep", "fprop", "bprop", "hprop")}
    # one lane tile: over several tiles a single hprop program would count the lane-independent residual
    # rows once per tile (the objective call counts them once)
    mem = emu_engine.createMemory(8)
    x["prog"].parameterize(trxfile.tiles_to_wide(mine.view("prog").parameters, parity.tile_arrays(ref, "params")[:1]), mem)
    x["prog"].run(ref.arrays["x"], mem)
    x["prep"].execute(mem)
    v = np.random.default_rng(5).standard_normal(ref.arrays["x"].size)
    two_sweeps = x["bprop"].run(x["fprop"].run(v, mem), mem)
    one_program = x["hprop"].run(v, mem)
    assert np.any(one_program != 0.0)
    parity.assert_close(one_program, two_sweeps, "J^T J v", rtol=1e-12)
