"""Collision shapes for the tape operators `collision_axes` and `collision_project_2`
(dexopt/src/ops.h:208-252,288-375).

Mirrors the reference's shape classes (tractor/include/tractor/collision/shape.h: CollisionSphereShape :51-63,
CollisionCylinderShape :139-171, ConvexPolyhedralCollisionShape :196-222) and ShapeCollisionPair
(collision/query.h:123-152): an object here owns a handle into the library's shape table; the handle is what a
tape carries as its uint64 constant where the reference carries the object's address.  The queries run on the
GPU (csrc/collision/trx_collide.h); there is no host implementation behind this module.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native

SPHERE, CYLINDER, POLYHEDRON = 1, 2, 3


class ShapeDesc(ctypes.Structure):
    """trx_shape_desc (include/trx_b200.h)"""
    _fields_ = [("kind", ctypes.c_int), ("center", ctypes.c_double * 3), ("radius", ctypes.c_double),
                ("length", ctypes.c_double), ("n_points", ctypes.c_uint32), ("n_planes", ctypes.c_uint32),
                ("points", ctypes.POINTER(ctypes.c_double)), ("planes", ctypes.POINTER(ctypes.c_double))]


def make_desc(kind, center=(0.0, 0.0, 0.0), radius=0.0, length=0.0, points=None, planes=None):
    """(desc, keepalive): the arrays must outlive the create call"""
    d = ShapeDesc()
    d.kind = kind
    d.center[:] = [float(v) for v in center]
    d.radius, d.length = float(radius), float(length)
    keep = []
    for name, arr, width in (("points", points, 3), ("planes", planes, 4)):
        if arr is None:
            setattr(d, "n_" + name, 0)
            continue
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1, width)
        keep.append(a)
        setattr(d, "n_" + name, a.shape[0])
        setattr(d, name, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return d, keep


class CollisionShape:
    def __init__(self, kind, **kw):
        L = _native.lib()
        desc, keep = make_desc(kind, **kw)
        h = ctypes.c_uint64()
        _native.check(L.trx_collision_shape_create(ctypes.byref(desc), ctypes.byref(h)))
        self.handle = h.value
        self.kind = kind

    @property
    def center(self):
        out = (ctypes.c_double * 3)()
        _native.check(_native.lib().trx_collision_shape_center(ctypes.c_uint64(self.handle), out))
        return np.array(out[:])

    def project(self, points, device=0):
        """CollisionShape::project for every row of `points`: (normals [n,3], distances [n])"""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty((pts.shape[0], 4))
        _native.check(_native.lib().trx_collision_project_points(
            ctypes.c_uint64(self.handle), pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(pts.shape[0]),
            out.ctypes.data_as(ctypes.c_void_p), device))
        return out[:, :3], out[:, 3]


def sphere(center, radius):
    return CollisionShape(SPHERE, center=center, radius=radius)


def cylinder(radius, length):
    return CollisionShape(CYLINDER, radius=radius, length=length)


def polyhedron(points, planes):
    return CollisionShape(POLYHEDRON, points=points, planes=planes)


class ShapeCollisionPair:
    def __init__(self, shape_a: CollisionShape, shape_b: CollisionShape):
        h = ctypes.c_uint64()
        _native.check(_native.lib().trx_collision_pair_create(ctypes.c_uint64(shape_a.handle),
                                                              ctypes.c_uint64(shape_b.handle), ctypes.byref(h)))
        self.handle = h.value
        self.shape_a, self.shape_b = shape_a, shape_b

    def update(self, poses_a, poses_b, device=0):
        """ShapeCollisionPair::update for every row (pose = translation xyz + quaternion xyzw):
        point_a [n,3], point_b [n,3], normal [n,3], distance [n]"""
        pa = np.ascontiguousarray(poses_a, dtype=np.float64).reshape(-1, 7)
        pb = np.ascontiguousarray(poses_b, dtype=np.float64).reshape(-1, 7)
        poses = np.ascontiguousarray(np.concatenate([pa, pb], 1))
        out = np.empty((pa.shape[0], 10))
        _native.check(_native.lib().trx_collision_query_pairs(
            ctypes.c_uint64(self.handle), poses.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(pa.shape[0]),
            out.ctypes.data_as(ctypes.c_void_p), device))
        return out[:, 0:3], out[:, 3:6], out[:, 6:9], out[:, 9]


def clear():
    _native.lib().trx_collision_clear()


def table():
    """the shape table as the kernels read it (records of csrc/collision/trx_collide.h)"""
    L = _native.lib()
    n = ctypes.c_uint64()
    _native.check(L.trx_collision_table(None, 0, ctypes.byref(n)))
    out = np.empty(n.value)
    _native.check(L.trx_collision_table(out.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
    return out
