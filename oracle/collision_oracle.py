"""TEST INFRASTRUCTURE -- never imported by the product (robio_2021_dexopt_b200/).

Independent computation of what `collision_axes` / `collision_project_2` report
(dexopt/src/ops.h:208-252,288-375 -> tractor/src/collision/query.cpp:43-154 -> Bullet btGjkEpaSolver3).

PARITY UNPINNED: Bullet is not in this image and the reference has no test at this boundary, so there is
nothing of the reference's to run or to compare against.  What can be pinned is the geometry itself, by a method
that shares nothing with the product's GJK / EPA (csrc/collision/trx_collide.h):

    distance(A, B)    = distance from the origin to  conv{a_i - b_j}          (origin outside)
    penetration(A, B) = min over the facets of conv{a_i - b_j} of their offset (origin inside)

with the convex hull of all vertex differences built by qhull (scipy.spatial.ConvexHull) and every facet visited.
Shapes follow tractor/include/tractor/collision/shape.h: sphere (:51-63) = centre + radius, handled as a point
with a margin; z-cylinder (:139-171) discretised into a 2 x n-gon prism (error <= r (1 - cos(pi / n)));
convex polyhedron (:196-222) = its points.  Result convention as btGjkEpa3.h: normal from B to A, distance
negative when overlapping.
"""
import numpy as np
from scipy.spatial import ConvexHull


def quat_rotate(q, v):
    """unit quaternion xyzw times vector (tractor/include/tractor/geometry/quaternion.h:71-104)"""
    x, y, z, w = q
    t = 2.0 * np.cross([x, y, z], v)
    return v + w * t + np.cross([x, y, z], t)


def pose_apply(pose, pts):
    pose = np.asarray(pose, float)
    return np.array([pose[:3] + quat_rotate(pose[3:], p) for p in np.atleast_2d(pts)])


def box_points(sx, sy, sz):
    h = 0.5 * np.array([sx, sy, sz])
    return np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], float) * h


def box_planes(sx, sy, sz):
    """nx ny nz offset with n.x + offset <= 0 inside (geometry/plane.h:24-26)"""
    out = []
    for axis, half in enumerate((0.5 * sx, 0.5 * sy, 0.5 * sz)):
        for sign in (-1.0, 1.0):
            n = [0.0, 0.0, 0.0]
            n[axis] = sign
            out.append(n + [-half])
    return np.array(out)


def cylinder_points(radius, length, n=2048):
    ang = 2.0 * np.pi * np.arange(n) / n
    ring = np.stack([radius * np.cos(ang), radius * np.sin(ang)], 1)
    return np.concatenate([np.c_[ring, np.full(n, 0.5 * length)], np.c_[ring, np.full(n, -0.5 * length)]])


def _closest_on_triangle(a, b, c):
    """closest point of triangle abc to the origin: interior projection or the best of the three edges"""
    best, best_d = None, np.inf
    n = np.cross(b - a, c - a)
    nn = n @ n
    if nn > 0:
        p = n * (a @ n) / nn  # projection of the origin onto the plane
        # inside test by barycentric coordinates
        m = np.stack([b - a, c - a], 1)
        uv, *_ = np.linalg.lstsq(m, p - a, rcond=None)
        if uv[0] >= 0 and uv[1] >= 0 and uv[0] + uv[1] <= 1:
            return p
    for s, e in ((a, b), (b, c), (c, a)):
        d = e - s
        t = 0.0 if d @ d == 0 else min(1.0, max(0.0, -(s @ d) / (d @ d)))
        q = s + t * d
        if q @ q < best_d:
            best, best_d = q, q @ q
    return best


def polytope_query(va, vb):
    """va, vb: world-frame vertex arrays.  Returns (distance, normal from B to A)."""
    diff = (va[:, None, :] - vb[None, :, :]).reshape(-1, 3)
    hull = ConvexHull(diff)
    eq = hull.equations  # n.x + off <= 0 inside
    if np.all(eq[:, 3] <= 0):
        # origin inside: minimum translation = nearest facet plane
        k = int(np.argmax(eq[:, 3]))
        return float(eq[k, 3]), -eq[k, :3].copy()
    best, best_d = None, np.inf
    for simplex in hull.simplices:
        q = _closest_on_triangle(*diff[simplex])
        if q @ q < best_d:
            best, best_d = q, q @ q
    d = float(np.sqrt(best_d))
    return d, best / d


def shape_vertices(shape, pose, n_cyl=2048):
    """(world vertices of the shape's core, margin)"""
    kind = shape["kind"]
    if kind == "sphere":
        return pose_apply(pose, [shape["center"]]), float(shape["radius"])
    if kind == "cylinder":
        return pose_apply(pose, cylinder_points(shape["radius"], shape["length"], n_cyl)), 0.0
    return pose_apply(pose, shape["points"]), 0.0


def pair_query(shape_a, pose_a, shape_b, pose_b, n_cyl=2048):
    va, ma = shape_vertices(shape_a, pose_a, n_cyl)
    vb, mb = shape_vertices(shape_b, pose_b, n_cyl)
    if len(va) == 1 and len(vb) == 1:
        v = va[0] - vb[0]
        d = float(np.linalg.norm(v))
        return d - ma - mb, v / d
    if len(va) * len(vb) < 4 or np.linalg.matrix_rank((va[:, None] - vb[None]).reshape(-1, 3) - (va[0] - vb[0])) < 3:
        raise ValueError("flat Minkowski difference: use a closed form")
    d, n = polytope_query(va, vb)
    return d - ma - mb, n


def point_cylinder(radius, length, p):
    """signed distance and normal (from the point to the shape, as project() reports it: query.h:111-121 with
    A = shape, B = point) for a z-cylinder, closed form"""
    p = np.asarray(p, float)
    rho = np.hypot(p[0], p[1])
    dr, dz = rho - radius, abs(p[2]) - 0.5 * length
    radial = np.array([p[0], p[1], 0.0]) / rho if rho > 0 else np.array([1.0, 0.0, 0.0])
    axial = np.array([0.0, 0.0, np.sign(p[2]) if p[2] != 0 else 1.0])
    if dr <= 0 and dz <= 0:  # inside: nearest surface
        if dr > dz:
            return dr, -radial
        return dz, -axial
    v = max(dr, 0.0) * radial + max(dz, 0.0) * axial  # from the closest surface point to p
    d = float(np.linalg.norm(v))
    return d, -v / d


def inside_polytope(planes, p, tol):
    planes = np.asarray(planes, float)
    return bool(np.all(planes[:, :3] @ p + planes[:, 3] <= tol))
