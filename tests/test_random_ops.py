"""In-tape random operators add_random_normal / add_random_uniform / dropout2 (dexopt/src/ops.h:120-204).

The reference draws from an unseeded thread_local std::mt19937, so there are no values to compare with: the tests
pin the generator (Philox4x32-10 known answers from the Random123 distribution), the statistics of the draws, that
every execute draws afresh, that a seed reproduces a run, and the derivative rules as the reference writes them
(prepare saves the mask; forward / reverse scale by it; random offsets pass the tangent through)."""
import ctypes

import numpy as np
import pytest

from robio_2021_dexopt_b200 import trxfile
from robio_2021_dexopt_b200.trxfile import Port

T8 = (64, 8)  # (bytes, lanes) of a Batch<double,8> slot
S1 = (8, 1)


def _arg(kind, is_input):
    size, lanes = kind
    return (size, 1 if is_input else 0, lanes, 0)


def _program(op_name, sig, n_inst=64):
    """n_inst instructions of one operator, every argument of instruction i in its own slot; inputs = the input
    arguments of instruction 0 broadcast by the caller (all instructions share the input slots), outputs = every
    output argument of every instruction"""
    kinds = [(T8 if tok[1] == "T" else S1, tok[0] == "i") for tok in sig.split()]
    ops = [(op_name, [_arg(k, i) for k, i in kinds])]
    addr = 0
    in_addr = []
    for k, is_in in kinds:
        if is_in:
            in_addr.append(addr)
            addr += k[0]
    inputs, outputs, instructions = [], [], []
    off = 0
    for (k, is_in), a in zip([x for x in kinds if x[1]], in_addr):
        inputs.append(Port(a, off, k[0], k[1], 0))
        off += k[0]
    off = 0
    for _ in range(n_inst):
        args, it = [], iter(in_addr)
        for k, is_in in kinds:
            if is_in:
                args.append(next(it))
            else:
                args.append(addr)
                outputs.append(Port(addr, off, k[0], k[1], 0))
                off += k[0]
                addr += k[0]
        instructions.append((0, args))
    return trxfile.encode_program(8, addr, ops, instructions, inputs=inputs, outputs=outputs), kinds


def _run(engine, blob, wide_inputs, lanes):
    x = engine.compile(blob)
    mem = engine.createMemory(lanes)
    return x, mem, x.run(np.asarray(wide_inputs, dtype=np.float64), mem)


def _engines():
    return [pytest.param("emu", id="emulation"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture
def engine(request, emu_engine):
    if request.param == "emu":
        return emu_engine
    import robio_2021_dexopt_b200 as pkg
    return pkg.CudaEngine(0)


def test_philox_known_answers(emu_engine):
    """Random123 kat_vectors: philox4x32-10"""
    f = emu_engine.lib.emu_philox4x32_10
    u32x4, u32x2 = ctypes.c_uint32 * 4, ctypes.c_uint32 * 2
    cases = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
              (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in cases:
        out = u32x4()
        f(u32x4(*ctr), u32x2(*key), out)
        assert tuple(out) == want


@pytest.mark.parametrize("engine", _engines(), indirect=True)
def test_add_random_normal_statistics(engine):
    lanes, n_inst = 64, 64
    blob, _ = _program("add_random_normal_8d", "iT1 iS1 oT1", n_inst)
    a = np.linspace(-1, 1, lanes)
    sigma = 0.25
    wide = np.concatenate([a, [sigma]])
    x, mem, out = _run(engine, blob, wide, lanes)
    draws = (out.reshape(n_inst, lanes) - a) / sigma
    assert abs(draws.mean()) < 5 / np.sqrt(draws.size) and abs(draws.std() - 1) < 0.05
    assert abs(np.mean(draws ** 3)) < 0.15 and abs(np.mean(draws ** 4) - 3) < 0.3
    # instructions and rollouts draw independently
    assert abs(np.corrcoef(draws[0], draws[1])[0, 1]) < 0.5 and len(np.unique(draws)) == draws.size
    # a second execute draws afresh
    again = x.run(wide, mem)
    assert not np.array_equal(again, out)


@pytest.mark.parametrize("engine", _engines(), indirect=True)
def test_add_random_uniform_statistics(engine):
    lanes, n_inst = 64, 64
    blob, _ = _program("add_random_uniform_8d", "iT1 iS1 iS1 oT1", n_inst)
    a = np.zeros(lanes)
    lo, hi = -0.01, 0.03
    _, _, out = _run(engine, blob, np.concatenate([a, [lo, hi]]), lanes)
    assert out.min() >= lo and out.max() < hi
    assert abs(out.mean() - 0.5 * (lo + hi)) < 5 * (hi - lo) / np.sqrt(12 * out.size)
    assert abs(out.var() - (hi - lo) ** 2 / 12) < 0.1 * (hi - lo) ** 2 / 12


@pytest.mark.parametrize("engine", _engines(), indirect=True)
def test_dropout2_mask_and_rate(engine):
    lanes, n_inst, rate = 64, 64, 0.3
    blob, _ = _program("dropout2_8d", "iT1 iS1 oT1 oT1", n_inst)
    a = np.linspace(1, 2, lanes)
    _, _, out = _run(engine, blob, np.concatenate([a, [rate]]), lanes)
    out = out.reshape(n_inst, 2, lanes)
    xs, ys = out[:, 0], out[:, 1]
    assert set(np.unique(ys)) == {0.0, 1.0 / (1.0 - rate)}
    assert np.array_equal(xs, a * ys)
    dropped = np.mean(ys == 0.0)
    assert abs(dropped - rate) < 5 * np.sqrt(rate * (1 - rate) / ys.size)


@pytest.mark.parametrize("engine", _engines(), indirect=True)
def test_seed_reproduces_a_run(request, engine, emu_engine):
    lanes = 16
    blob, _ = _program("add_random_normal_8d", "iT1 iS1 oT1", 4)
    wide = np.concatenate([np.zeros(lanes), [1.0]])
    if engine is emu_engine:
        set_seed = lambda s: emu_engine.lib.emu_set_random_seed(ctypes.c_ulonglong(s))  # noqa: E731
    else:
        from robio_2021_dexopt_b200 import _native
        _native.lib().trx_set_random_seed.argtypes = [ctypes.c_uint64]
        set_seed = lambda s: _native.lib().trx_set_random_seed(s)  # noqa: E731
    set_seed(42)
    x, mem, first = _run(engine, blob, wide, lanes)
    second = x.run(wide, mem)
    set_seed(42)
    assert np.array_equal(x.run(wide, mem), first) and np.array_equal(x.run(wide, mem), second)
    set_seed(43)
    assert not np.array_equal(x.run(wide, mem), first)


@pytest.mark.parametrize("engine", _engines(), indirect=True)
def test_derivative_rules_as_written(engine):
    """prepare_dropout2 saves the mask, forward / reverse scale by it and give the rate no derivative;
    forward / reverse of the random offsets pass the tangent through (dexopt/src/ops.h:137-143,186-198)"""
    lanes = 8
    rng = np.random.default_rng(3)
    # prepare: (a, b, x, y) -> s = y
    blob, _ = _program("prepare_dropout2_8d", "iT1 iS1 iT1 iT1 oT1", 1)
    a, xx, y = rng.standard_normal(lanes), rng.standard_normal(lanes), rng.standard_normal(lanes)
    _, _, s = _run(engine, blob, np.concatenate([a, [0.3], xx, y]), lanes)  # wide layout: ports in declaration order
    assert np.array_equal(s, y)
    # forward: (s, da, db) -> dx = da * s, dy = 0
    blob, _ = _program("forward_dropout2_8d", "iT1 iT1 iS1 oT1 oT1", 1)
    da = rng.standard_normal(lanes)
    _, _, out = _run(engine, blob, np.concatenate([y, da, [7.0]]), lanes)
    assert np.array_equal(out[:lanes], da * y) and np.all(out[lanes:] == 0)
    # reverse: (s, dx, dy) -> da = dx * s, db = 0
    blob, _ = _program("reverse_dropout2_8d", "iT1 oT1 oS1 iT1 iT1", 1)
    dx = rng.standard_normal(lanes)
    _, _, out = _run(engine, blob, np.concatenate([y, dx, np.ones(lanes)]), lanes)
    assert out[lanes] == 0.0 and np.array_equal(out[:lanes], dx * y)
    blob, _ = _program("reverse_add_random_normal_8d", "oT1 oS1 iT1", 1)
    _, _, out = _run(engine, blob, dx, lanes)
    assert out[lanes] == 0.0 and np.array_equal(out[:lanes], dx)
    blob, _ = _program("forward_add_random_uniform_8d", "iT1 iS1 iS1 oT1", 1)
    _, _, out = _run(engine, blob, np.concatenate([da, [1.0, 2.0]]), lanes)
    assert np.array_equal(out, da)
