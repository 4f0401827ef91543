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
