"""Shapes and query wrappers for the collision tests: the same descriptions registered with the TEST-ONLY host
emulation (tests/emu/vm_emu.cpp, CPU suite) or with the product library (GPU suite), and handed to the oracle."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import collision_oracle as co  # noqa: E402

from robio_2021_dexopt_b200.collision import CYLINDER, POLYHEDRON, SPHERE, make_desc  # noqa: E402


def tetra_points():
    return np.array([[0.03, 0.0, -0.01], [-0.02, 0.025, -0.012], [-0.015, -0.03, -0.008], [0.002, 0.001, 0.035]])


def planes_of(points):
    """outward planes nx ny nz offset of the convex hull of `points` (qhull)"""
    from scipy.spatial import ConvexHull
    eq = ConvexHull(points).equations
    # one plane per distinct facet orientation
    uniq = []
    for e in eq:
        if not any(np.allclose(e, u, atol=1e-12) for u in uniq):
            uniq.append(e)
    return np.array(uniq)


SHAPES = {
    "sphere": dict(kind="sphere", center=[0.004, -0.003, 0.002], radius=0.02),
    "ball": dict(kind="sphere", center=[0.0, 0.0, 0.0], radius=0.035),
    "cylinder": dict(kind="cylinder", radius=0.04, length=0.06),  # hand8.xacro:24-26
    "box": dict(kind="polyhedron", points=co.box_points(0.06, 0.06, 0.06), planes=co.box_planes(0.06, 0.06, 0.06)),
    "slab": dict(kind="polyhedron", points=co.box_points(0.2, 0.15, 0.02), planes=co.box_planes(0.2, 0.15, 0.02)),
    "tetra": dict(kind="polyhedron", points=tetra_points(), planes=planes_of(tetra_points())),
}


def desc_of(shape):
    if shape["kind"] == "sphere":
        return make_desc(SPHERE, center=shape["center"], radius=shape["radius"])
    if shape["kind"] == "cylinder":
        return make_desc(CYLINDER, radius=shape["radius"], length=shape["length"])
    return make_desc(POLYHEDRON, points=shape["points"], planes=shape["planes"])


class EmuCollision:
    """the emulator's copy of trx_collision_* (same record builder, same query code, on the host)"""

    def __init__(self, lib):
        self.lib = lib
        c = ctypes
        lib.emu_collision_shape_create.argtypes = [c.c_void_p, c.POINTER(c.c_uint64)]
        lib.emu_collision_pair_create.argtypes = [c.c_uint64, c.c_uint64, c.POINTER(c.c_uint64)]
        lib.emu_collision_query_pairs.argtypes = [c.c_uint64, c.c_void_p, c.c_uint64, c.c_void_p]
        lib.emu_collision_project_points.argtypes = [c.c_uint64, c.c_void_p, c.c_uint64, c.c_void_p]
        lib.emu_last_error.restype = c.c_char_p

    def shape(self, shape):
        desc, keep = desc_of(shape)
        h = ctypes.c_uint64()
        st = self.lib.emu_collision_shape_create(ctypes.byref(desc), ctypes.byref(h))
        if st != 0:
            raise RuntimeError("emu status %d: %s" % (st, self.lib.emu_last_error().decode()))
        return h.value

    def pair(self, a, b):
        h = ctypes.c_uint64()
        st = self.lib.emu_collision_pair_create(a, b, ctypes.byref(h))
        if st != 0:
            raise RuntimeError("emu status %d: %s" % (st, self.lib.emu_last_error().decode()))
        return h.value

    def query(self, pair, poses_a, poses_b):
        poses = np.ascontiguousarray(np.concatenate([np.reshape(poses_a, (-1, 7)), np.reshape(poses_b, (-1, 7))], 1))
        out = np.empty((poses.shape[0], 12))
        assert self.lib.emu_collision_query_pairs(pair, poses.ctypes.data_as(ctypes.c_void_p), poses.shape[0],
                                                  out.ctypes.data_as(ctypes.c_void_p)) == 0
        return out

    def project(self, shape, points):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty((pts.shape[0], 4))
        assert self.lib.emu_collision_project_points(shape, pts.ctypes.data_as(ctypes.c_void_p), pts.shape[0],
                                                     out.ctypes.data_as(ctypes.c_void_p)) == 0
        return out


def random_pose(rng, spread):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return np.concatenate([rng.uniform(-spread, spread, 3), q])


def random_pose_pairs(seed, n, spread):
    rng = np.random.default_rng(seed)
    return (np.array([random_pose(rng, spread) for _ in range(n)]),
            np.array([random_pose(rng, spread) for _ in range(n)]))


def in_shape(shape, pose, p, tol):
    """is world point p inside / on the posed shape"""
    pose = np.asarray(pose, float)
    qi = np.concatenate([-pose[3:6], pose[6:]])
    local = co.quat_rotate(qi, np.asarray(p) - pose[:3])
    if shape["kind"] == "sphere":
        return np.linalg.norm(local - shape["center"]) <= shape["radius"] + tol
    if shape["kind"] == "cylinder":
        return np.hypot(local[0], local[1]) <= shape["radius"] + tol and abs(local[2]) <= 0.5 * shape["length"] + tol
    return co.inside_polytope(shape["planes"], local, tol)
