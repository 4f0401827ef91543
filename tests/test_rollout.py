"""The re-rolled rollout objective (csrc/rollout, csrc/trx_rollout.cu) against the reference.

Expected values: the reference SimpleEngine on the reference-recorded tape of the same problem
(tests/golden/rollout_*.trxb, values only).  CPU tests run the frame phases through the TEST-ONLY host
driver (tests/emu/rollout_emu.cpp); GPU tests call the product through the C ABI
(trx_objective_create_rollout / _evaluate / _residuals / _gd_steps)."""
import os

import numpy as np
import pytest

import parity
import rollout_util as ru

FIXTURE_NAMES = sorted(ru.FIXTURES)


@pytest.fixture(scope="module")
def emu():
    return ru.RolloutEmu()


# ------------------------------------------------------------------ CPU: the phases themselves

@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_frame_phases_match_reference(emu, golden, name):
    """residual vector, loss and J^T r of the hand-written forward / adjoint frame equal the reference's tape"""
    b = golden(name)
    opts = ru.FIXTURES[name]
    ns = emu.info(opts, "n_scalar")
    pw, want_r = ru.wide_from_tiles(b, ns)
    r, loss, g = emu.evaluate(opts, pw, b.arrays["x"])
    parity.assert_close(r, want_r, "residuals")
    want = float(b.arrays["loss_total"][0])
    assert abs(loss - want) <= 1e-11 * abs(want)
    parity.assert_close(g, b.arrays["g_total"], "gradient")


def test_checkpoint_is_the_carried_state_only(emu):
    """what crosses a frame boundary (and HBM): joint positions, commands, body state, tracked link poses"""
    opts = ru.FIXTURES["rollout_handarm_h64.trxb"]
    assert emu.info(opts, "checkpoint_doubles") == 29 + 29 + 13 + 7 * 7
    # independent of the horizon
    assert emu.info(dict(opts, frames=256), "checkpoint_doubles") == emu.info(opts, "checkpoint_doubles")
    assert emu.info(dict(opts, frames=256), "region_doubles") == emu.info(opts, "region_doubles")


def test_adjoint_with_given_seeds_is_linear(emu, golden):
    """J^T v for a given v (the Gauss-Newton building block): linear in v and equal to J^T r for v = r"""
    b = golden("rollout_hand8_linear.trxb")
    opts = ru.FIXTURES["rollout_hand8_linear.trxb"]
    ns = emu.info(opts, "n_scalar")
    pw, _ = ru.wide_from_tiles(b, ns)
    pw = pw[:, :3].copy()
    x = b.arrays["x"]
    r, _, g = emu.evaluate(opts, pw, x)
    _, _, g_r = emu.evaluate(opts, pw, x, want_r=False, seed=r)
    parity.assert_close(g_r, g, "J^T r through explicit seeds")
    rng = np.random.default_rng(3)
    v = rng.standard_normal(r.size)
    _, _, g_v = emu.evaluate(opts, pw, x, want_r=False, seed=v)
    _, _, g_2 = emu.evaluate(opts, pw, x, want_r=False, seed=2.0 * v + r)
    parity.assert_close(g_2, 2.0 * g_v + g, "linearity", rtol=1e-9, atol=1e-12)


# ------------------------------------------------------------------ GPU: the product

def _objective(cuda_engine, opts, lanes, **kw):
    import robio_2021_dexopt_b200 as pkg
    return pkg.RolloutObjective(cuda_engine, lanes, **dict(opts, **kw))


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_rollout_objective_matches_reference(cuda_engine, golden, name):
    b = golden(name)
    opts = ru.FIXTURES[name]
    n_tiles = len(parity.tile_arrays(b, "params"))
    obj = _objective(cuda_engine, opts, 8 * n_tiles)
    ns = obj.info("n_scalar_rows")
    pw, want_r = ru.wide_from_tiles(b, ns)
    obj.parameterize(pw)
    x = b.arrays["x"]
    r = obj.residuals(x)
    parity.assert_close(r, want_r, "residuals")
    loss, g = obj.evaluate(x)
    want = float(b.arrays["loss_total"][0])
    assert abs(loss - want) <= 1e-11 * abs(want), (loss, want)
    parity.assert_close(g, b.arrays["g_total"], "gradient")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rollout_handarm_h64.trxb", "rollout_hand8_linear.trxb"])
def test_rollout_gd_iterates_match_reference(cuda_engine, golden, name):
    """ten GradientDescentSolver steps resident on the device (trx_objective_gd_steps) against the reference's
    x_10 and its ten losses (solvers/gd.h:43-129; lr 0.01, momentum 0, dexopt.cpp:199-202)"""
    import robio_2021_dexopt_b200 as pkg
    b = golden(name)
    opts = ru.FIXTURES[name]
    obj = _objective(cuda_engine, opts, 16)
    pw, _ = ru.wide_from_tiles(b, obj.info("n_scalar_rows"))
    obj.parameterize(pw)
    solver = pkg.GradientDescentSolver(obj, learning_rate=float(b.arrays["gd/learning_rate"][0]))
    solver.gather(b.arrays["x"])
    losses = solver.steps_on_device(10)
    want = b.arrays["gd/loss"]
    assert np.all(np.abs(losses - want) <= 1e-9 * np.abs(want)), (losses, want)
    parity.assert_close(solver.scatter(), b.arrays["gd/x"], "x after 10 GD steps", rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_rollout_partial_group_and_chunks(cuda_engine, emu, golden):
    """rollout counts that do not fill a CTA's group, and rollouts walked in chunks under a byte budget,
    against the host execution of the same phases (which the CPU tests pin to the reference)"""
    name = "rollout_handarm_long.trxb"
    b = golden(name)
    opts = ru.FIXTURES[name]
    x = b.arrays["x"]
    rng = np.random.default_rng(11)
    lanes = 21
    pw = np.stack([rng.uniform(-0.01, 0.01, lanes), rng.uniform(-0.01, 0.01, lanes), rng.uniform(0, 2 * np.pi, lanes)])
    want_r, want_loss, want_g = emu.evaluate(opts, pw, x)
    for budget in (0, 8 * emu.info(opts, "checkpoint_doubles") * 8 * opts["frames"]):
        obj = _objective(cuda_engine, opts, lanes, max_device_bytes=budget)
        assert (obj.info("n_chunks") > 1) == (budget > 0)
        obj.parameterize(pw)
        parity.assert_close(obj.residuals(x), want_r, "residuals")
        loss, g = obj.evaluate(x)
        assert abs(loss - want_loss) <= 1e-11 * abs(want_loss)
        parity.assert_close(g, want_g, "gradient")


@pytest.mark.gpu
@pytest.mark.parametrize("checkpoint", ["state", "policy", "full"])
def test_rollout_checkpoint_modes_match_reference(cuda_engine, golden, checkpoint):
    """the three checkpoint contents (TRX_OBJECTIVE_CHECKPOINT_*: the adjoint sweep recomputes the whole frame,
    kinematics and contacts only, or nothing) against the reference's loss and gradient of the 24-frame rollout"""
    name = "rollout_handarm_long.trxb"
    b = golden(name)
    opts = ru.FIXTURES[name]
    n_tiles = len(parity.tile_arrays(b, "params"))
    obj = _objective(cuda_engine, opts, 8 * n_tiles, checkpoint=checkpoint)
    assert obj.info("checkpoint_mode") == {"state": 0, "policy": 1, "full": 2}[checkpoint]
    carried = obj.info("carried_state_bytes_per_unit")
    if checkpoint == "state":
        assert obj.info("checkpoint_bytes_per_unit") < carried + 128  # rounded up to whole 128-byte lines
    pw, _ = ru.wide_from_tiles(b, obj.info("n_scalar_rows"))
    obj.parameterize(pw)
    loss, g = obj.evaluate(b.arrays["x"])
    want = float(b.arrays["loss_total"][0])
    assert abs(loss - want) <= 1e-11 * abs(want), (loss, want)
    parity.assert_close(g, b.arrays["g_total"], "gradient")


@pytest.mark.gpu
def test_rollout_agrees_with_tape_objective_at_the_benchmark_shape(cuda_engine):
    """BASELINE configs[2] shape (41 -> 64 -> 83, 29 controlled joints), 16 frames x 64 rollouts: the re-rolled
    kernels against the interpreted flat tape of the same problem (itself pinned to the reference)"""
    import robio_2021_dexopt_b200 as pkg
    from robio_2021_dexopt_b200 import trxfile
    opts = dict(kind="handarm", hidden=64, frames=16)
    lanes = 64
    bundle = pkg.build_dexsyn(fuse_dense=1, programs="prog,prep,bprop", **opts)
    tape = pkg.Objective(cuda_engine, bundle, lanes)
    roll = _objective(cuda_engine, opts, lanes)
    rng = np.random.default_rng(5)
    pw = np.stack([rng.uniform(-0.01, 0.01, lanes), rng.uniform(-0.01, 0.01, lanes), rng.uniform(0, 2 * np.pi, lanes)])
    tape.parameterize(pw)
    roll.parameterize(pw)
    x = bundle.arrays["x"]
    l0, g0 = tape.evaluate(x)
    l1, g1 = roll.evaluate(x)
    assert abs(l0 - l1) <= 1e-11 * abs(l0), (l0, l1)
    parity.assert_close(g1, g0, "gradient", rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_rollout_full_size_properties(cuda_engine):
    """BASELINE configs[2] at full size (horizon 256 x 1024 rollouts): deterministic, and additive over
    rollout shards with the lane-independent rows counted once"""
    opts = dict(kind="handarm", hidden=64, frames=256)
    lanes = 1024
    rng = np.random.default_rng(9)
    pw = np.stack([rng.uniform(-0.01, 0.01, lanes), rng.uniform(-0.01, 0.01, lanes), rng.uniform(0, 2 * np.pi, lanes)])
    import robio_2021_dexopt_b200 as pkg
    x = pkg.build_dexsyn(frames=1, programs="prog", kind="handarm", hidden=64).arrays["x"]
    full = _objective(cuda_engine, opts, lanes)
    full.parameterize(pw)
    l_a, g_a = full.evaluate(x)
    l_b, g_b = full.evaluate(x)
    assert l_a == l_b and np.array_equal(g_a, g_b)
    assert np.isfinite(l_a) and np.all(np.isfinite(g_a))
    # device memory is rollouts x frames x checkpoint (carried state + the frame's policy values), nothing else
    assert full.info("device_bytes") < 1024 * 256 * full.info("checkpoint_bytes_per_unit") + 20e6
    # default checkpoint: carried state (120 doubles) + policy values (240) + link poses, policy input, contact
    # points and velocities, body centre (382), in whole 128-byte lines
    assert full.info("checkpoint_bytes_per_unit") <= 8 * (120 + 240 + 382 + 15)
    half0 = _objective(cuda_engine, opts, lanes // 2)
    half1 = _objective(cuda_engine, opts, lanes // 2, scalar_rows=False)
    half0.parameterize(pw[:, :lanes // 2].copy())
    half1.parameterize(pw[:, lanes // 2:].copy())
    l0, g0 = half0.evaluate(x)
    l1, g1 = half1.evaluate(x)
    assert abs((l0 + l1) - l_a) <= 1e-11 * abs(l_a)
    parity.assert_close(g0 + g1, g_a, "sharded gradient", rtol=1e-10, atol=1e-12)
