"""CPU suite: the product's lowering (levels, bundles, port maps) and operator bodies, run through
the TEST-ONLY host emulation, against golden vectors computed by the reference SimpleEngine."""
import pytest

import parity


@pytest.mark.parametrize("fixture", ["optests_8d.trxb", "optests_d.trxb"])
def test_every_operator_matches_reference(emu_engine, golden, fixture):
    bundle = golden(fixture)
    names = parity.op_names(bundle)
    assert len(names) >= 70
    for name, grad in names:
        parity.check_single_op(emu_engine, bundle, name, grad == "grad")


def test_fk_chain_two_tiles(emu_engine, golden):
    parity.check_problem(emu_engine, golden("fkchain24.trxb"))


def test_fk_chain_one_tile(emu_engine, golden):
    parity.check_problem(emu_engine, golden("fkchain24.trxb"), n_tiles=1)


@pytest.mark.parametrize("fixture", ["dexsyn_small.trxb.xz", "dexsyn_dropout_outer2.trxb.xz"])
def test_dexsyn_rollout(emu_engine, golden, fixture):
    parity.check_problem(emu_engine, golden(fixture))


@pytest.mark.parametrize("fixture", ["fkchain24.trxb", "dexsyn_small.trxb.xz", "dexsyn_dropout_outer2.trxb.xz"])
@pytest.mark.parametrize("chain", [1, 3])
def test_accumulate_folding_of_the_adjoint_tape(emu_engine, golden, fixture, chain, monkeypatch):
    """`rop -> temp; move/add temp -> g` lowered as one record writing g (trx_lower.h, fold_accumulate):
    same gradient as the reference, fewer records"""
    bundle = golden(fixture)
    monkeypatch.setenv("TRX_CHAIN", str(chain))
    plain = emu_engine.compile(bundle.programs["bprop"])

    def compile_folded(blob):
        monkeypatch.setenv("TRX_FOLD", "1")
        try:
            return emu_engine.compile(blob)
        finally:
            monkeypatch.delenv("TRX_FOLD")

    folded = compile_folded(bundle.programs["bprop"])
    assert folded.stat("n_folded") > 0
    assert folded.stat("n_bundles") < plain.stat("n_bundles")
    parity.check_problem(emu_engine, bundle, check_tangent=False, compile_bprop=compile_folded)


@pytest.mark.parametrize("chain", [1, 2, 4, 8, 64])
def test_chain_levelisation_keeps_the_numbers(emu_engine, golden, chain, monkeypatch):
    """instructions appended to their producer's chain (trx_lower.h) run in program order on the same
    threads: residuals, gradient and tangent stay the reference's for every chain cap, levels shrink"""
    bundle = golden("dexsyn_small.trxb.xz")
    monkeypatch.setenv("TRX_CHAIN", "1")
    flat = emu_engine.compile(bundle.programs["bprop"]).stat("n_levels")
    monkeypatch.setenv("TRX_CHAIN", str(chain))
    x = emu_engine.compile(bundle.programs["bprop"])
    assert (x.stat("n_levels") < flat) if chain > 1 else (x.stat("n_levels") == flat)
    assert (x.stat("n_chained") > 0) == (chain > 1)
    parity.check_problem(emu_engine, bundle)


@pytest.mark.parametrize("fixture", ["optests_8d.trxb", "fkchain24.trxb", "dexsyn_small.trxb.xz"])
def test_stream_layout_invariants(emu_engine, golden, fixture):
    """what the staged kernel relies on: page-aligned per-warp streams, no record across a page boundary,
    the same number of BARRIER words in every warp's stream"""
    bundle = golden(fixture)
    for name, blob in bundle.programs.items():
        emu_engine.compile(blob).check_streams()


@pytest.mark.parametrize("fixture", ["optests_8d.trxb", "optests_d.trxb", "dexsyn_small.trxb.xz", "dexsyn_dropout_outer2.trxb.xz"])
def test_adjoint_reads_forward_operands_in_place(emu_engine, golden, fixture):
    """objective lowering step trx_direct.h: bprop arguments that name x_prep's operand copies are redirected
    to the forward slots, reverse_mul becomes reverse_x_muld, the now useless prepare instructions go away;
    the gradient stays the reference's"""
    bundle = golden(fixture)
    if fixture.startswith("optests"):
        n = 0
        for name, has_grad in ((l[0], l[1] == "grad") for l in parity.op_names(bundle)):
            if not has_grad:
                continue
            pre = "op/%s/" % name
            sub = type("B", (), {"programs": {"prep": bundle.programs[pre + "prep"], "bprop": bundle.programs[pre + "bprop"]}})
            prep, bprop, stats = emu_engine.direct_prepare(sub)
            n += stats["prepare_dropped"]
            x_prog = emu_engine.compile(bundle.programs[pre + "prog"])
            mem = emu_engine.createMemory(8)
            x_prog.run(bundle.arrays[pre + "x"], mem)
            emu_engine.compile(prep).execute(mem)
            gx = emu_engine.compile(bprop).run(bundle.arrays[pre + "gy"], mem)
            parity.assert_close(gx, bundle.arrays[pre + "gx"], name + " reverse, operands in place")
        assert n >= 10  # mul, the vec3 / quat / twist products, gate, ...
        return
    prep, bprop, stats = emu_engine.direct_prepare(bundle)
    assert stats["prepare_dropped"] > 0 and stats["arguments_redirected"] + stats["mul_rewritten"] > 0
    assert emu_engine.compile(prep).stat("n_instructions") < emu_engine.compile(bundle.programs["prep"]).stat("n_instructions")
    parity.check_problem(emu_engine, bundle, check_tangent=False, blobs={"prep": prep, "bprop": bprop})


def test_level_schedule_is_shallower_than_the_tape(emu_engine, golden):
    bundle = golden("dexsyn_small.trxb.xz")
    x = emu_engine.compile(bundle.programs["prog"])
    assert x.stat("n_levels") < x.stat("n_instructions") / 10
    prep = emu_engine.compile(bundle.programs["prep"])
    assert prep.stat("n_levels") == 1  # every prepare instruction is independent


@pytest.mark.parametrize("fixture", ["fkchain24.trxb", "dexsyn_small.trxb.xz"])
def test_gradient_descent_iterates(emu_engine, golden, fixture):
    import numpy as np
    bundle = golden(fixture)
    obj = parity.ExecutableObjective(emu_engine, bundle, 2)
    lr = float(bundle.arrays["gd/learning_rate"][0])
    state = {"x": bundle.arrays["x"].copy()}

    def step():
        loss, g = obj.evaluate(state["x"])
        state["x"] = state["x"] + (-lr * g)
        return loss, state["x"]

    parity.check_gd_iterates(step, bundle)


def _fused_objective(emu_engine, bundle, ref, ring_slots=192, sync_penalty=900):
    from robio_2021_dexopt_b200 import trxfile
    fwd, adj = emu_engine.compile_fused(bundle, ring_slots, sync_penalty)
    prog = bundle.view("prog")
    params = parity.tile_arrays(ref, "params")
    mem = emu_engine.createMemory(8 * len(params))
    if prog.parameters:
        fwd.parameterize(trxfile.tiles_to_wide(prog.parameters, params), mem)
    r = fwd.run(ref.arrays["x"], mem)
    parity.assert_close(r, trxfile.tiles_to_wide(prog.outputs, parity.tile_arrays(ref, "r")), "residuals")
    g = adj.run(r, mem)
    parity.assert_close(g, ref.arrays["g_total"], "gradient")
    return fwd, adj


@pytest.mark.parametrize("fixture", ["fkchain24.trxb", "dexsyn_small.trxb.xz", "dexsyn_dropout_outer2.trxb.xz"])
@pytest.mark.parametrize("ring_slots,sync_penalty", [(192, 900), (16, 100), (192, 0)])
def test_fused_lowering_of_reference_tapes(emu_engine, golden, fixture, ring_slots, sync_penalty):
    """the objective's latency-oriented lowering (chain scheduling, per-warp slot rings, write-through
    only where needed; csrc/trx_fuse.h) on tapes recorded by the reference"""
    ref = golden(fixture)
    fwd, adj = _fused_objective(emu_engine, ref, ref, ring_slots, sync_penalty)
    assert fwd.stat("n_ring_reads") > 0 and adj.stat("n_ring_reads") > 0


@pytest.mark.parametrize("fuse_dense", [0, 1])
def test_fused_lowering_of_product_tapes(emu_engine, golden, fuse_dense):
    import robio_2021_dexopt_b200 as pkg
    ref = golden("dexsyn_dropout_outer2.trxb.xz")
    mine = pkg.build_dexsyn(frames=2, hidden=4, extra_fixed=2, dropout=1, outer=2, fuse_dense=fuse_dense)
    fwd, adj = _fused_objective(emu_engine, mine, ref)
    assert fwd.stat("n_ring_only_writes") > 0  # some results never leave the chip
