"""Constraint operators (tractor/include/tractor/core/constraints.h:10-265, geometry/ops.h:998-1130; SURVEY.md
section 8 row N4): every registered operator whose name contains "constraint", in every mode the reference has --
compute / prepare / forward / reverse and the modes buildConstraints (src/core/gradients.cpp:764-997) collects into
SolverBase's project / barrier_init / barrier_step / barrier_diagonal programs (solvers/base.h:124-137), plus the
penalty rules -- each as a single-instruction program against the reference SimpleEngine's output
(tests/golden/constraint_ops_{8d,d}.trxb, `ref_tool constraint_tests`).  These rules are comparisons, sums and one
division: bit-exact on the host emulation; on the device u*u + v*v contracts to a fused multiply-add, so the bar
there is 1e-14 relative (infinities and signs exact)."""
import numpy as np
import pytest

# (the reference indexes six components of three-component arguments in barrier_init / barrier_step of the quaternion
# trust region and barrier_step of the vector3 one, geometry/ops.h:1076-1081,1106-1125: what it touches past the end is
# undefined -- the fixture pads every argument -- and the components the arguments have are compared)


def same(y, want, exact):
    if exact:
        nan = np.isnan(y) & np.isnan(want)
        return np.array_equal(y[~nan], want[~nan])
    return np.allclose(y, want, rtol=1e-14, atol=0.0, equal_nan=True)


def check_constraint_ops(engine, bundle, exact):
    names = [line for line in bundle.texts["ops"].splitlines() if line.strip()]
    assert len(names) >= 50
    modes = set()
    for name in names:
        pre = "op/%s/" % name
        x = engine.compile(bundle.programs[pre + "prog"])
        mem = engine.createMemory(8)
        y = x.run(bundle.arrays[pre + "x"], mem)
        want = bundle.arrays[pre + "y"]
        assert y.shape == want.shape, name
        assert same(y, want, exact), (name, y, want)
        modes.add(name.split("_")[0])
    return names, modes


@pytest.mark.parametrize("fixture", ["constraint_ops_8d.trxb", "constraint_ops_d.trxb"])
def test_constraint_operators_on_the_emulated_vm(emu_engine, golden, fixture):
    names, modes = check_constraint_ops(emu_engine, golden(fixture), True)
    assert {"project", "barrier", "penalty", "forward", "reverse"} <= modes


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["constraint_ops_8d.trxb", "constraint_ops_d.trxb"])
def test_constraint_operators_on_the_device(cuda_engine, golden, fixture):
    check_constraint_ops(cuda_engine, golden(fixture), False)


def check_constraint_programs(engine, bundle, exact):
    """tests/golden/constrained_small.trxb (`ref_tool constrained`): a bounded least-squares tape whose variables
    carry range / trust-region constraints (as robot/jointtypes.h:105-119 records them); the reference's
    buildConstraints derived project / barrier_init / barrier_step / barrier_diagonal, the reference engine ran them
    as the interior-point solver does (solvers/ip.h:212-246,287-293,475-478).  The device runs the same four
    programs on one Memory after x_prog (the constraint arguments live in the tape's memory)."""
    import parity
    from robio_2021_dexopt_b200 import trxfile
    x_prog = engine.compile(bundle.programs["prog"])
    mem = engine.createMemory(8)
    prog = bundle.view("prog")
    x_prog.parameterize(trxfile.tiles_to_wide(prog.parameters, [bundle.arrays["tile0/params"]]), mem)
    r = x_prog.run(bundle.arrays["x"], mem)
    parity.assert_close(r, trxfile.tiles_to_wide(prog.outputs, [bundle.arrays["tile0/r"]]), "residuals")
    x_proj = engine.compile(bundle.programs["project"])
    x_proj.parameterize(bundle.arrays["padding"], mem)
    y = x_proj.run(bundle.arrays["step"], mem)
    assert same(y, bundle.arrays["project/y"], exact), (y, bundle.arrays["project/y"])
    # the projected step keeps every constrained variable inside its padded range: not the identity
    assert np.any(y != bundle.arrays["step"])
    y = engine.compile(bundle.programs["barrier_init"]).run(bundle.arrays["step2"], mem)
    assert same(y, bundle.arrays["barrier_init/y"], exact)
    y = engine.compile(bundle.programs["barrier_step"]).run(bundle.arrays["step"], mem)
    assert same(y, bundle.arrays["barrier_step/y"], exact)
    x_diag = engine.compile(bundle.programs["barrier_diagonal"])
    x_diag.execute(mem)
    y = x_diag.output(mem)
    assert same(y, bundle.arrays["barrier_diagonal/y"], exact)
    assert np.count_nonzero(y) >= 6


def test_constraint_programs_on_the_emulated_vm(emu_engine, golden):
    check_constraint_programs(emu_engine, golden("constrained_small.trxb"), True)


@pytest.mark.gpu
def test_constraint_programs_on_the_device(cuda_engine, golden):
    check_constraint_programs(cuda_engine, golden("constrained_small.trxb"), False)
