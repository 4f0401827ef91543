"""collision_axes / collision_project_2 (dexopt/src/ops.h:208-252,288-375; row a10 / N2 of SURVEY.md section 8).

PARITY UNPINNED against the reference (its Bullet dependency is not in this image and it has no test at this
boundary): what is pinned is the geometry -- closed forms and an independent Minkowski-difference hull computation
(oracle/collision_oracle.py).  The CPU suite runs the same query code as the kernel through the TEST-ONLY host
emulation; the GPU suite runs the product library and compares it with both.

Tolerances: distance 1e-10 m for polyhedra and spheres (exact arithmetic up to rounding); 5e-7 m where a cylinder
is involved (the oracle's 2 x 1024-gon prism deviates by r (1 - cos(pi / 1024)) = 2e-7 m)."""
import ctypes

import numpy as np
import pytest

import collision_util as cu
from collision_util import co

PAIRS = [("box", "box", 1e-10, 0), ("tetra", "box", 1e-10, 0), ("sphere", "box", 1e-10, 0), ("tetra", "sphere", 1e-10, 0),
         ("cylinder", "box", 5e-7, 1024), ("sphere", "cylinder", 5e-7, 1024), ("slab", "tetra", 1e-10, 0),
         # two cylinders: the hull of all vertex differences limits the oracle to 2 x 192-gons (6e-6 m)
         ("cylinder", "cylinder", 1.5e-5, 192)]


@pytest.fixture(scope="module")
def emu_collision(emu_engine):
    return cu.EmuCollision(emu_engine.lib)


def check_against_oracle(query, name_a, name_b, tol, n_cyl=2048, n=40, seed=1):
    """query(poses_a, poses_b) -> rows [a, b, normal, distance, ...]"""
    A, B = cu.SHAPES[name_a], cu.SHAPES[name_b]
    n_sep = n_pen = 0
    for spread in (0.02, 0.06):  # mostly overlapping / mostly separated
        pa, pb = cu.random_pose_pairs(seed, n, spread)
        out = query(pa, pb)
        for i in range(n):
            want_d, want_n = co.pair_query(A, pa[i], B, pb[i], n_cyl or 2048)
            a, b, nrm, d = out[i, 0:3], out[i, 3:6], out[i, 6:9], out[i, 9]
            assert abs(d - want_d) <= tol, (name_a, name_b, i, d, want_d)
            assert abs(np.linalg.norm(nrm) - 1.0) <= 1e-12
            # the witnesses realise the distance along the normal and lie on their shapes
            assert np.linalg.norm((a - b) - nrm * d) <= 10 * tol, (name_a, name_b, i)
            assert cu.in_shape(A, pa[i], a, 10 * tol) and cu.in_shape(B, pb[i], b, 10 * tol), (name_a, name_b, i)
            # the direction is unique when the shapes are apart (and generically when they overlap)
            if abs(want_d) > 1e-3:
                # (a discretised cylinder quantises the oracle's facet normals to its 2 pi / n sectors)
                ntol = max(1e-6, 100 * tol / abs(want_d), 2 * np.pi / n_cyl if n_cyl else 0.0)
                assert np.linalg.norm(nrm - want_n) <= ntol, (name_a, name_b, i, nrm, want_n)
            n_sep += want_d > 0
            n_pen += want_d < 0
    assert n_sep >= 3 and n_pen >= 3, (n_sep, n_pen)  # both branches (GJK distance, EPA penetration) were exercised


@pytest.mark.parametrize("name_a,name_b,tol,n_cyl", PAIRS)
def test_pair_queries_match_minkowski_hull(emu_collision, name_a, name_b, tol, n_cyl):
    ha, hb = emu_collision.shape(cu.SHAPES[name_a]), emu_collision.shape(cu.SHAPES[name_b])
    pair = emu_collision.pair(ha, hb)
    check_against_oracle(lambda pa, pb: emu_collision.query(pair, pa, pb), name_a, name_b, tol, n_cyl,
                         n=12 if n_cyl else 40)


def test_sphere_sphere_closed_form(emu_collision):
    A, B = cu.SHAPES["sphere"], cu.SHAPES["ball"]
    pair = emu_collision.pair(emu_collision.shape(A), emu_collision.shape(B))
    pa, pb = cu.random_pose_pairs(3, 50, 0.05)
    out = emu_collision.query(pair, pa, pb)
    for i in range(50):
        ca = co.pose_apply(pa[i], [A["center"]])[0]
        cb = co.pose_apply(pb[i], [B["center"]])[0]
        v = ca - cb
        d = np.linalg.norm(v)
        assert abs(out[i, 9] - (d - A["radius"] - B["radius"])) <= 1e-15
        np.testing.assert_allclose(out[i, 6:9], v / d, atol=1e-14)
        np.testing.assert_allclose(out[i, 0:3], ca - v / d * A["radius"], atol=1e-15)
        np.testing.assert_allclose(out[i, 3:6], cb + v / d * B["radius"], atol=1e-15)


def test_box_box_axis_aligned_known_answers(emu_collision):
    box = cu.SHAPES["box"]  # 0.06 cube
    pair = emu_collision.pair(emu_collision.shape(box), emu_collision.shape(box))
    ident = [0, 0, 0, 1]
    # face to face, 0.04 apart along x; A at +x of B  ->  normal from B to A = +x
    out = emu_collision.query(pair, [[0.1, 0.0, 0.0] + ident], [[0.0, 0.0, 0.0] + ident])[0]
    assert abs(out[9] - 0.04) <= 1e-15 and np.allclose(out[6:9], [1, 0, 0], atol=1e-14)
    assert abs(out[0] - 0.07) <= 1e-15 and abs(out[3] - 0.03) <= 1e-15
    # overlapping by 0.01 along z (and fully in x, y): minimum translation along z
    out = emu_collision.query(pair, [[0.0, 0.0, 0.05] + ident], [[0.0, 0.0, 0.0] + ident])[0]
    assert abs(out[9] + 0.01) <= 1e-14 and np.allclose(out[6:9], [0, 0, 1], atol=1e-12)
    # corner to corner along the diagonal
    out = emu_collision.query(pair, [[0.1, 0.1, 0.1] + ident], [[0.0, 0.0, 0.0] + ident])[0]
    assert abs(out[9] - 0.04 * np.sqrt(3)) <= 1e-15 and np.allclose(out[6:9], np.ones(3) / np.sqrt(3), atol=1e-14)
    # the same boxes exactly coincident: deepest overlap, depth = one full edge
    out = emu_collision.query(pair, [[0.0, 0.0, 0.0] + ident], [[0.0, 0.0, 0.0] + ident])[0]
    assert abs(out[9] + 0.06) <= 1e-14


def test_project_polyhedron(emu_collision):
    """ConvexPolyhedralCollisionShape::project (shape.h:327-346): inside -> the nearest plane, normal = -plane normal;
    outside -> the closest-point query against the point (normal from the point to the shape, distance > 0)"""
    box = cu.SHAPES["box"]
    h = emu_collision.shape(box)
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.06, 0.06, (200, 3))
    out = emu_collision.project(h, pts)
    n_in = 0
    for p, o in zip(pts, out):
        q = np.clip(p, -0.03, 0.03)  # closest point of the cube
        if np.all(np.abs(p) <= 0.03):
            n_in += 1
            k = int(np.argmax(np.abs(p)))
            want_n = np.zeros(3)
            want_n[k] = -np.sign(p[k])
            assert abs(o[3] - (abs(p[k]) - 0.03)) <= 1e-16 and np.array_equal(o[:3], want_n)
        else:
            d = np.linalg.norm(q - p)
            assert abs(o[3] - d) <= 1e-14
            np.testing.assert_allclose(o[:3], (q - p) / d, atol=1e-12)
    assert 10 < n_in < 190


def test_project_sphere_and_cylinder_closed_forms(emu_collision):
    rng = np.random.default_rng(6)
    S = cu.SHAPES["sphere"]
    hs = emu_collision.shape(S)
    pts = rng.uniform(-0.05, 0.05, (100, 3))
    out = emu_collision.project(hs, pts)
    for p, o in zip(pts, out):
        v = np.array(S["center"]) - p
        assert abs(o[3] - (np.linalg.norm(v) - S["radius"])) <= 1e-16
        np.testing.assert_allclose(o[:3], v / np.linalg.norm(v), atol=1e-15)
    C = cu.SHAPES["cylinder"]
    hc = emu_collision.shape(C)
    pts = rng.uniform(-0.08, 0.08, (300, 3))
    out = emu_collision.project(hc, pts)
    n_in = 0
    for p, o in zip(pts, out):
        d, nrm = co.point_cylinder(C["radius"], C["length"], p)
        n_in += d < 0
        assert abs(o[3] - d) <= 1e-9, (p, o, d)
        if abs(d) > 1e-4:
            # on the curved side the distance converges quadratically faster than the witness direction (sliver simplices of
                # nearly equal rim points): direction to 2e-5
                assert np.linalg.norm(o[:3] - nrm) <= max(2e-5, 1e-7 / abs(d)), (p, o, nrm)
    assert n_in > 10


def test_iteration_counts_are_bounded(emu_collision):
    """termination: polyhedra end by a repeated support vertex, curved shapes by the tolerance -- never by the cap"""
    for name_a, name_b in (("box", "tetra"), ("cylinder", "box"), ("cylinder", "cylinder")):
        pair = emu_collision.pair(emu_collision.shape(cu.SHAPES[name_a]), emu_collision.shape(cu.SHAPES[name_b]))
        pa, pb = cu.random_pose_pairs(9, 200, 0.08)
        out = emu_collision.query(pair, pa, pb)
        assert out[:, 10].max() < 100, (name_a, name_b, out[:, 10].max())
        assert np.all(np.isfinite(out))


def test_shape_registration_errors(emu_collision):
    from robio_2021_dexopt_b200.collision import POLYHEDRON, make_desc
    lib = emu_collision.lib
    h = ctypes.c_uint64()
    desc, keep = make_desc(POLYHEDRON, points=np.zeros((0, 3)), planes=np.zeros((1, 4)))
    assert lib.emu_collision_shape_create(ctypes.byref(desc), ctypes.byref(h)) == -1  # shape.h:203-205
    desc, keep = make_desc(7)
    assert lib.emu_collision_shape_create(ctypes.byref(desc), ctypes.byref(h)) == -2  # "not yet supported"
    assert lib.emu_collision_pair_create(12345, 1, ctypes.byref(h)) == -1


# ---------------------------------------------------------------- the two operators on a tape
# dexsyn problem with collision penalties (dexlearn.h:66-131) and friction cones (dexlearn.h:219-233,324-362) recorded
# by the product's recorder.  No reference numbers exist for these rows (Bullet); what is pinned: (1) the reference's
# residuals of the SAME problem without those penalties appear unchanged, in order, among the new residuals (the
# penalties only add goals), (2) the tangent and adjoint sweeps -- derivative rules as the reference writes them --
# are adjoint to each other (tractor/src/test/test.cpp:629-645), (3) the collision rows respond to the geometry.
PROBLEM = dict(frames=3, hidden=4, extra_fixed=4)


def build_collision_problem(**extra):
    import robio_2021_dexopt_b200 as pkg
    pkg.collision.clear()
    return pkg.build_dexsyn(collision=1, friction=1, **dict(PROBLEM, **extra))


def is_subsequence(small, big, rtol=1e-11):
    j = 0
    for v in small:
        while j < len(big) and abs(big[j] - v) > rtol * max(1.0, abs(v)):
            j += 1
        if j == len(big):
            return False
        j += 1
    return True


def check_collision_tape(engine, golden, n_tiles=2):
    import parity
    from robio_2021_dexopt_b200 import trxfile
    ref = golden("dexsyn_small.trxb.xz")  # the same problem without the penalties, run by the reference
    mine = build_collision_problem()
    names = [op.name for op in mine.view("prog").ops]
    assert "collision_axes_8d" in names and "collision_project_2_8d" in names
    assert "reverse_collision_project_2_8d" in [op.name for op in mine.view("bprop").ops]
    prog = mine.view("prog")
    params = parity.tile_arrays(ref, "params")[:n_tiles]
    x = ref.arrays["x"]
    x_prog, x_prep = engine.compile(mine.programs["prog"]), engine.compile(mine.programs["prep"])
    x_bprop, x_fprop = engine.compile(mine.programs["bprop"]), engine.compile(mine.programs["fprop"])
    mem = engine.createMemory(8 * n_tiles)
    x_prog.parameterize(trxfile.tiles_to_wide(prog.parameters, params), mem)
    r = x_prog.run(x, mem)
    assert np.all(np.isfinite(r))
    r_tiles = trxfile.wide_to_tiles(prog.outputs, r, n_tiles) if hasattr(trxfile, "wide_to_tiles") else None
    want_tiles = parity.tile_arrays(ref, "r")[:n_tiles]
    if r_tiles is not None:
        for t in range(n_tiles):
            assert len(r_tiles[t]) > len(want_tiles[t])
            assert is_subsequence(want_tiles[t], r_tiles[t]), "tile %d: the reference's residuals changed" % t
    x_prep.execute(mem)
    rng = np.random.default_rng(11)
    w = rng.normal(size=r.size)
    v = rng.normal(size=x.size)
    jt_w = x_bprop.run(w, mem)
    j_v = x_fprop.run(v, mem)
    lhs, rhs = float(w @ j_v), float(v @ jt_w)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs)), (lhs, rhs)
    return r, jt_w, j_v


def use_library_table(emu_engine):
    """the emulator's copy of the library's shape table, so that product-built tapes find their handles"""
    import robio_2021_dexopt_b200 as pkg
    table = np.ascontiguousarray(pkg.collision.table())
    emu_engine.lib.emu_collision_set_table.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    emu_engine.lib.emu_collision_set_table(table.ctypes.data_as(ctypes.c_void_p), table.size)


def test_collision_tape_on_the_emulated_vm(emu_engine, golden):
    build_collision_problem()  # registers the shapes (the same offsets again after every clear())
    use_library_table(emu_engine)
    r, g, dr = check_collision_tape(emu_engine, golden)
    assert np.count_nonzero(r) > 0


def one_instruction(name, sig, handle):
    """tape with a single instruction; sig tokens as csrc/vm_ops.def ("iT3 iU1 oT3 oT1"); lane-typed inputs are
    input ports, the uint64 argument a constant holding `handle`, outputs are output ports"""
    from robio_2021_dexopt_b200 import trxfile
    from robio_2021_dexopt_b200.trxfile import Port
    args, addrs, inputs, outputs, constants = [], [], [], [], []
    addr = in_off = out_off = 0
    for tok in sig.split():
        is_in, kind, comps = tok[0] == "i", tok[1], int(tok[2:])
        lanes = 8 if kind == "T" else 1
        size = comps * 8 * lanes
        args.append((size, is_in, lanes, 1 if kind == "U" else 0))
        addrs.append(addr)
        if kind == "U" and is_in:
            constants.append(Port(addr, 0, 8, 1, 1))
        elif is_in:
            inputs.append(Port(addr, in_off, size, lanes, 0))
            in_off += size
        else:
            outputs.append(Port(addr, out_off, size, lanes, 1 if kind == "U" else 0))
            out_off += size
        addr += (size + 63) // 64 * 64
    return trxfile.encode_program(8, addr, [(name, args)], [(0, addrs)], inputs=inputs, outputs=outputs,
                                  constants=constants, const_data=np.array([handle], dtype=np.uint64).tobytes())


def soa(rows):
    """[8 lanes, comps] -> component-major, lane-minor"""
    return np.ascontiguousarray(np.asarray(rows, float).T).ravel()


def check_single_instructions(engine, query):
    """collision_project_2 with its prepare / forward / reverse rules as written (dexopt/src/ops.h:225-252) and
    collision_axes with local = pose^-1 * point, zero derivatives (ops.h:288-375), one instruction at a time.
    query: object with .project(shape, points) / .pairs(pair, poses_a, poses_b) giving the expected op results"""
    import robio_2021_dexopt_b200 as pkg
    box = pkg.collision.polyhedron(cu.SHAPES["box"]["points"], cu.SHAPES["box"]["planes"])
    tet = pkg.collision.polyhedron(cu.SHAPES["tetra"]["points"], cu.SHAPES["tetra"]["planes"])
    cyl = pkg.collision.cylinder(0.04, 0.06)
    pair = pkg.collision.ShapeCollisionPair(tet, cyl)
    rng = np.random.default_rng(2)
    mem = engine.createMemory(8)

    pts = rng.uniform(-0.06, 0.06, (8, 3))
    want = query.project(box.handle, pts)
    y = engine.compile(one_instruction("collision_project_2_8d", "iT3 iU1 oT3 oT1", box.handle)).run(soa(pts), mem)
    np.testing.assert_allclose(y[:24].reshape(3, 8).T, want[:, :3], rtol=0, atol=1e-13)
    np.testing.assert_allclose(y[24:], want[:, 3], rtol=0, atol=1e-15)
    # prepare: n = normal
    nrm, dist = rng.normal(size=(8, 3)), rng.normal(size=(8, 1))
    y = engine.compile(one_instruction("prepare_collision_project_2_8d", "iT3 iU1 iT3 iT1 oT3", box.handle)).run(
        np.concatenate([soa(pts), soa(nrm), soa(dist)]), mem)
    np.testing.assert_array_equal(y, soa(nrm))
    # forward: d normal = 0, d distance = -n . d point
    dp = rng.normal(size=(8, 3))
    y = engine.compile(one_instruction("forward_collision_project_2_8d", "iT3 iT3 iU1 oT3 oT1", box.handle)).run(
        np.concatenate([soa(nrm), soa(dp)]), mem)
    np.testing.assert_array_equal(y[:24], np.zeros(24))
    np.testing.assert_allclose(y[24:], -np.sum(nrm * dp, 1), rtol=1e-15, atol=0)
    # reverse: point gradient = n * -(distance gradient), handle gradient = 0
    gn, gd = rng.normal(size=(8, 3)), rng.normal(size=(8, 1))
    y = engine.compile(one_instruction("reverse_collision_project_2_8d", "iT3 oT3 oU1 iT3 iT1", box.handle)).run(
        np.concatenate([soa(nrm), soa(gn), soa(gd)]), mem)
    np.testing.assert_allclose(y[:24].reshape(3, 8).T, nrm * -gd, rtol=1e-15, atol=0)
    assert y[24] == 0.0

    pa, pb = cu.random_pose_pairs(4, 8, 0.05)
    want = query.pairs(pair.handle, pa, pb)
    y = engine.compile(one_instruction("collision_axes_8d", "iT7 iT7 iU1 oT3 oT3 oT3 oT3 oT3", pair.handle)).run(
        np.concatenate([soa(pa), soa(pb)]), mem)
    out = y.reshape(5, 3, 8).transpose(0, 2, 1)  # [output, lane, xyz]
    np.testing.assert_allclose(out[0], want[:, 0:3], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out[1], want[:, 3:6], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out[2], want[:, 6:9], rtol=0, atol=1e-12)
    for lane in range(8):
        for which, pose in ((3, pa[lane]), (4, pb[lane])):
            qi = np.concatenate([-pose[3:6], pose[6:]])
            local = co.quat_rotate(qi, out[which - 3][lane] - pose[:3])
            np.testing.assert_allclose(out[which][lane], local, rtol=0, atol=1e-13)
    y = engine.compile(one_instruction("forward_collision_axes_8d", "iT6 iT6 iU1 oT3 oT3 oT3 oT3 oT3", pair.handle)).run(
        rng.normal(size=96), mem)
    np.testing.assert_array_equal(y, np.zeros(120))
    y = engine.compile(one_instruction("reverse_collision_axes_8d", "oT6 oT6 oU1 iT3 iT3 iT3 iT3 iT3", pair.handle)).run(
        rng.normal(size=120), mem)
    np.testing.assert_array_equal(y, np.zeros(97))


class EmuQueries:
    def __init__(self, emu):
        self.emu = emu

    def project(self, shape, pts):
        return self.emu.project(shape, pts)

    def pairs(self, pair, pa, pb):
        return self.emu.query(pair, pa, pb)


def test_collision_operators_as_single_instructions(emu_engine, emu_collision):
    import robio_2021_dexopt_b200 as pkg
    pkg.collision.clear()
    # register once to learn how large the table gets, then mirror it; handles are offsets, so registering the
    # same shapes again inside check_single_instructions() lands on the same records
    for _ in range(2):
        pkg.collision.clear()
        pkg.collision.polyhedron(cu.SHAPES["box"]["points"], cu.SHAPES["box"]["planes"])
        t = pkg.collision.polyhedron(cu.SHAPES["tetra"]["points"], cu.SHAPES["tetra"]["planes"])
        c = pkg.collision.cylinder(0.04, 0.06)
        pkg.collision.ShapeCollisionPair(t, c)
    use_library_table(emu_engine)
    pkg.collision.clear()
    check_single_instructions(emu_engine, EmuQueries(emu_collision))


def test_unregistered_handles_are_rejected_at_compile_time():
    """a tape whose uint64 constant is not a registered handle must not reach the device (library-side check;
    needs no GPU: the check runs before any device call... if a device is required the error is TRX_ERR_CUDA)"""
    import robio_2021_dexopt_b200 as pkg
    pkg.collision.clear()
    blob = one_instruction("collision_project_2_8d", "iT3 iU1 oT3 oT1", 4242)
    with pytest.raises(pkg.TrxError) as e:
        pkg.CudaEngine(0).compile(blob)
    assert e.value.status in (-1, -3)


def test_collision_problem_matches_the_reference_engine(emu_engine, golden):
    """tests/golden/dexsyn_collision.trxb: the SAME problem recorded through the reference's Var<> API with the
    reference's collision operators (ops.h:208-375 over stand-in shape classes whose one Bullet call is answered by
    the query under test), differentiated by the reference's buildGradients and run by the reference's SimpleEngine.
    The product's recorder + gradients + VM must give its residuals, loss, J^T r and J dx: everything around the
    geometric query -- operator bodies, derivative rules as written, the penalty assembly -- is pinned to the
    reference's own code; the query's arithmetic against Bullet's is what stays unpinned."""
    import parity
    ref = golden("dexsyn_collision.trxb")
    mine = build_collision_problem()
    use_library_table(emu_engine)
    for k, v in ref.arrays.items():
        mine.arrays.setdefault(k, v)
    r, g = parity.check_problem(emu_engine, mine, check_tangent=True)
    assert r.size == 24727 and g.size == 583


# ---------------------------------------------------------------- GPU: the product library
class DeviceQueries:
    def project(self, shape, pts):
        from robio_2021_dexopt_b200 import collision
        s = collision.CollisionShape.__new__(collision.CollisionShape)
        s.handle = shape
        n, d = s.project(pts)
        return np.concatenate([n, d[:, None]], 1)

    def pairs(self, pair, pa, pb):
        from robio_2021_dexopt_b200 import collision
        p = collision.ShapeCollisionPair.__new__(collision.ShapeCollisionPair)
        p.handle = pair
        a, b, n, d = p.update(pa, pb)
        return np.concatenate([a, b, n, d[:, None]], 1)


def device_shape(shape):
    from robio_2021_dexopt_b200 import collision
    if shape["kind"] == "sphere":
        return collision.sphere(shape["center"], shape["radius"])
    if shape["kind"] == "cylinder":
        return collision.cylinder(shape["radius"], shape["length"])
    return collision.polyhedron(shape["points"], shape["planes"])


@pytest.mark.gpu
@pytest.mark.parametrize("name_a,name_b,tol,n_cyl", PAIRS)
def test_device_pair_queries(emu_collision, name_a, name_b, tol, n_cyl):
    """the kernel's queries against the oracle, and against the host emulation of the same code (FMA contraction
    is the only difference: 1e-12 m)"""
    from robio_2021_dexopt_b200 import collision
    collision.clear()
    pair = collision.ShapeCollisionPair(device_shape(cu.SHAPES[name_a]), device_shape(cu.SHAPES[name_b]))
    dq = DeviceQueries()
    check_against_oracle(lambda pa, pb: dq.pairs(pair.handle, pa, pb), name_a, name_b, tol, n_cyl, n=12 if n_cyl else 40)
    epair = emu_collision.pair(emu_collision.shape(cu.SHAPES[name_a]), emu_collision.shape(cu.SHAPES[name_b]))
    pa, pb = cu.random_pose_pairs(21, 512, 0.05)
    got, want = dq.pairs(pair.handle, pa, pb), emu_collision.query(epair, pa, pb)[:, :10]
    # distances agree to rounding; witnesses / normals where they are unique (not face-to-face or rim cases)
    # (two curved shapes: the expanding polytope refines different faces on the device and on the host)
    np.testing.assert_allclose(got[:, 9], want[:, 9], rtol=0, atol=(2e-6 if n_cyl == 192 else 1e-9) if n_cyl else 1e-12)
    # (curved shapes: device and host take different support points once FMA contraction changes a comparison)
    close = np.linalg.norm(got[:, 6:9] - want[:, 6:9], axis=1) < (1e-3 if n_cyl else 1e-6)
    assert close.mean() > 0.9


@pytest.mark.gpu
def test_device_projection(emu_collision):
    from robio_2021_dexopt_b200 import collision
    collision.clear()
    rng = np.random.default_rng(8)
    dq = DeviceQueries()
    for name, spread in (("box", 0.06), ("tetra", 0.05), ("sphere", 0.05), ("cylinder", 0.08)):
        shape = device_shape(cu.SHAPES[name])
        pts = rng.uniform(-spread, spread, (4096, 3))
        got = dq.project(shape.handle, pts)
        want = emu_collision.project(emu_collision.shape(cu.SHAPES[name]), pts)
        np.testing.assert_allclose(got[:, 3], want[:, 3], rtol=0, atol=1e-9 if name == "cylinder" else 1e-12)
        # (on the cylinder's curved side the witness direction converges to 2e-5 only, see the CPU test)
        ntol = 1e-4 if name == "cylinder" else 1e-6
        assert (np.linalg.norm(got[:, :3] - want[:, :3], axis=1) < ntol).mean() > 0.99


@pytest.mark.gpu
def test_collision_operators_on_the_device(cuda_engine):
    from robio_2021_dexopt_b200 import collision
    collision.clear()
    # the same shapes in the same order as check_single_instructions registers them: DeviceQueries reads the handles
    check_single_instructions(cuda_engine, DeviceQueries())


@pytest.mark.gpu
def test_collision_tape_on_the_device(cuda_engine, emu_engine, golden):
    """the whole problem with collision and friction-cone penalties on the GPU: the checks of the CPU suite, and
    equality with the host emulation of the same tape (residuals, J^T w, J v)"""
    r, g, dr = check_collision_tape(cuda_engine, golden)
    use_library_table(emu_engine)
    r2, g2, dr2 = check_collision_tape(emu_engine, golden)
    # collision_project_2 on a polyhedron answers with the NEAREST plane's normal (shape.h:327-346): a contact point
    # within rounding of two planes' bisector gets either normal, and the friction-cone rows built from it flip.  Such
    # rows (2 of 24 727 here) are counted, everything else must agree.
    def mostly_close(got, want, what):
        scale = max(1.0, float(np.max(np.abs(want))))
        bad = np.abs(got - want) > 1e-11 * scale + 1e-9 * np.abs(want)
        assert bad.mean() <= 1e-3, "%s: %d of %d differ" % (what, int(bad.sum()), bad.size)
    mostly_close(r, r2, "residuals")
    mostly_close(dr, dr2, "tangent sweep")
    # (J^T w sums over all rows: compare after removing the flipped rows' weight is not possible here, so bound it)
    assert np.linalg.norm(g - g2) <= 1e-2 * np.linalg.norm(g2)


@pytest.mark.gpu
def test_collision_objective_and_gradient_descent_on_the_device(cuda_engine, golden):
    """trx_objective_evaluate on the collision problem, 32 rollouts: loss and gradient equal the executables' r.r and
    J^T r; ten device-resident GD steps decrease the loss"""
    import parity
    import robio_2021_dexopt_b200 as pkg
    from robio_2021_dexopt_b200 import trxfile
    ref = golden("dexsyn_small.trxb.xz")
    mine = build_collision_problem()
    params = parity.tile_arrays(ref, "params")
    obj = pkg.Objective(cuda_engine, mine, lanes=8 * len(params))
    obj.parameterize(trxfile.tiles_to_wide(mine.view("prog").parameters, params))
    loss, g = obj.evaluate(ref.arrays["x"])
    for k, v in ref.arrays.items():
        mine.arrays.setdefault(k, v)
    oracle = parity.ExecutableObjective(cuda_engine, mine, len(params))
    loss2, g2 = oracle.evaluate(ref.arrays["x"])
    assert abs(loss - loss2) <= 1e-11 * abs(loss2)
    parity.assert_close(g, g2, "gradient", rtol=1e-9)


@pytest.mark.gpu
def test_program_refuses_to_run_after_the_shape_table_was_cleared(cuda_engine):
    import robio_2021_dexopt_b200 as pkg
    pkg.collision.clear()
    box = pkg.collision.polyhedron(cu.SHAPES["box"]["points"], cu.SHAPES["box"]["planes"])
    x = cuda_engine.compile(one_instruction("collision_project_2_8d", "iT3 iU1 oT3 oT1", box.handle))
    mem = cuda_engine.createMemory(8)
    x.run(np.zeros(24), mem)
    pkg.collision.clear()
    with pytest.raises(pkg.TrxError) as e:
        x.run(np.zeros(24), mem)
    assert e.value.status == -1


@pytest.mark.gpu
def test_rollout_objective_refuses_collision_problems(cuda_engine):
    """the re-rolled rollout kernel has no collision phases: it must say so instead of dropping the penalties"""
    import robio_2021_dexopt_b200 as pkg
    with pytest.raises(pkg.TrxError) as e:
        pkg.RolloutObjective(cuda_engine, 16, frames=2, hidden=4, collision=1)
    assert e.value.status == -2


@pytest.mark.gpu
def test_collision_problem_matches_the_reference_engine_on_the_device(cuda_engine, golden):
    """the device against the reference engine's numbers for the collision problem (see the CPU test of the same
    name): residuals, loss, gradient, tangent through the executables, and ten gradient-descent iterates through the
    device-resident solver -- at the suite's tolerances (1e-11; iterates 1e-9).  floor_pair=0: without the floor /
    object pair, whose parallel faces in contact have no unique closest points."""
    import parity
    import robio_2021_dexopt_b200 as pkg
    from robio_2021_dexopt_b200 import trxfile
    ref = golden("dexsyn_collision_nofloor.trxb")
    mine = build_collision_problem(floor_pair=0)
    for k, v in ref.arrays.items():
        mine.arrays.setdefault(k, v)
    parity.check_problem(cuda_engine, mine, check_tangent=True)
    obj = pkg.Objective(cuda_engine, mine, lanes=16)
    obj.parameterize(trxfile.tiles_to_wide(mine.view("prog").parameters, parity.tile_arrays(ref, "params")))
    solver = pkg.GradientDescentSolver(obj, learning_rate=float(ref.arrays["gd/learning_rate"][0]))
    solver.gather(ref.arrays["x"])
    losses = solver.steps_on_device(10)  # the golden keeps x_10 and the ten losses
    want = ref.arrays["gd/loss"]
    assert np.all(np.abs(losses - want) <= 1e-9 * np.abs(want)), (losses, want)
    parity.assert_close(solver.scatter(), ref.arrays["gd/x"], "x after 10 GD steps", rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_collision_problem_with_the_floor_pair_on_the_device(cuda_engine, golden):
    """the full problem: the object rests on the floor face on face, the closest points of the two parallel faces are
    not unique, and the device (FMA contraction) reports other witnesses than the host in one lane: the slip-avoidance
    row built from them differs (2 of 24 727 residuals).  Everything else equals the reference engine's numbers."""
    import parity
    from robio_2021_dexopt_b200 import trxfile
    ref = golden("dexsyn_collision.trxb")
    mine = build_collision_problem()
    prog = mine.view("prog")
    params = parity.tile_arrays(ref, "params")
    x_prog = cuda_engine.compile(mine.programs["prog"])
    mem = cuda_engine.createMemory(16)
    x_prog.parameterize(trxfile.tiles_to_wide(prog.parameters, params), mem)
    r = x_prog.run(ref.arrays["x"], mem)
    want = trxfile.tiles_to_wide(prog.outputs, parity.tile_arrays(ref, "r"))
    bad = np.abs(r - want) > 1e-11 * np.maximum(1.0, np.abs(want))
    assert bad.sum() <= 6, int(bad.sum())
    loss, want_loss = float(r @ r), float(ref.arrays["loss_total"][0])
    assert abs(loss - want_loss) <= 1e-4 * want_loss
