// trx_collide.h -- closest points / penetration between convex shapes, per lane, on the device.
//
// Stands for what the reference computes on the HOST inside two tape operators
//   collision_axes      dexopt/src/ops.h:288-375  -> ShapeCollisionPair::update   collision/query.h:128-141
//   collision_project_2 dexopt/src/ops.h:208-252  -> CollisionShape::project      collision/query.h:111-121,
//                                                    ConvexPolyhedralCollisionShape::project  shape.h:327-346
// both of which end in doCollisionQuery (src/collision/query.cpp:43-154): Bullet's btGjkEpaSolver3_Distance and,
// when that reports an overlap, btGjkEpaSolver3_Penetration, over the support mappings of shape.h:51-227
// (sphere :60-63, z-cylinder :160-171, convex polyhedron :214-222) posed by CollisionShapeSupport
// (query.h:43-72).  Result convention (btGjkEpa3.h, restated): separated -> witnesses = the closest points,
// normal = (a - b) / |a - b| (from B to A), distance = |a - b| > 0;  overlapping -> the minimum translation:
// normal = -(outward normal of the closest face of A (-) B), distance = -depth, b = a - n_face * depth.
//
// PARITY UNPINNED: Bullet is not in this image and the reference holds no test at this boundary (SURVEY.md
// section 8c).  This file computes the geometric quantities themselves -- Gilbert-Johnson-Keerthi distance on the
// Minkowski difference, expanding-polytope penetration, both restated from their published algorithms -- in fp64
// with tight tolerances (Bullet: GJK_ACCURACY / EPA_ACCURACY 1e-4 in single-, 1e-12 in double-precision builds, 128 vertices); tests check them against closed forms
// and an independent Minkowski-difference hull computation (oracle/collision_oracle.py).  Spheres are treated
// as a point core with a margin, which is exact, where Bullet iterates on the support mapping.
//
// Shapes live in one table of doubles (the "shape heap") that the host registers (trx_collision_shape_create /
// trx_collision_pair_create) and the library mirrors to the device; the uint64 constant an instruction carries
// where the reference carries a host pointer is the record's offset in that table.
#pragma once
#include <stdint.h>

#include "../vm_math.cuh"

#ifdef __CUDACC__
#define TRX_COL_FN __host__ __device__ __noinline__ inline
#else
#define TRX_COL_FN inline
#endif

namespace trxcol {
using trxvm::cross;
using trxvm::dot;
using trxvm::PoseV;
using trxvm::Q4;
using trxvm::qinv;
using trxvm::qrot;
using trxvm::V3;

enum ShapeKind { SHAPE_SPHERE = 1, SHAPE_CYLINDER = 2, SHAPE_POLYHEDRON = 3, SHAPE_PAIR = 16 };
// record layout, doubles from the record's offset
enum { REC_KIND = 0, REC_CENTER = 1, REC_RADIUS = 4, REC_LENGTH = 5, REC_NPOINTS = 6, REC_NPLANES = 7, REC_DATA = 8 };
enum { PAIR_A = 1, PAIR_B = 2, PAIR_SIZE = 3 };

struct Shape {
  const double *rec;
  int kind, n_points, n_planes;
  TRX_HD const double *points() const { return rec + REC_DATA; }
  TRX_HD const double *planes() const { return rec + REC_DATA + 3 * n_points; }
  TRX_HD V3 center() const { return V3{rec[REC_CENTER], rec[REC_CENTER + 1], rec[REC_CENTER + 2]}; }
  TRX_HD double radius() const { return rec[REC_RADIUS]; }
  TRX_HD double length() const { return rec[REC_LENGTH]; }
  // a sphere is its centre (core) inflated by its radius (margin)
  TRX_HD double margin() const { return kind == SHAPE_SPHERE ? rec[REC_RADIUS] : 0.0; }
};
TRX_HD Shape shapeAt(const double *heap, uint64_t handle) {
  Shape s;
  s.rec = heap + handle;
  s.kind = (int)s.rec[REC_KIND];
  s.n_points = (int)s.rec[REC_NPOINTS];
  s.n_planes = (int)s.rec[REC_NPLANES];
  return s;
}

struct Result {
  V3 a, b, n;  // witness on A, witness on B, normal from B to A (world frame)
  double d;    // signed distance
  int status;  // 0 separated, 1 overlapping, 2 degenerate (no direction could be found)
  int gjk_iterations, epa_iterations;
};

TRX_HD double norm2(const V3 &a) { return dot(a, a); }

// support mapping of the shape's CORE in its own frame (shape.h:60-63 without the radius, :160-171, :214-222)
TRX_HD V3 supportLocal(const Shape &s, const V3 &d) {
  if (s.kind == SHAPE_SPHERE) return s.center();
  if (s.kind == SHAPE_CYLINDER) {
    const double sign = d.z < 0.0 ? -1.0 : (d.z > 0.0 ? 1.0 : 0.0);
    const double an = sqrt(d.x * d.x + d.y * d.y);
    V3 p = V3{0.0, 0.0, sign * (s.length() * 0.5)};
    if (an != 0.0) {
      const double f = s.radius() / an;
      p.x = d.x * f;
      p.y = d.y * f;
    }
    return p;
  }
  const double *pts = s.points();
  int best = 0;
  double best_dot = pts[0] * d.x + pts[1] * d.y + pts[2] * d.z;
  for (int i = 1; i < s.n_points; i++) {
    const double v = pts[3 * i] * d.x + pts[3 * i + 1] * d.y + pts[3 * i + 2] * d.z;
    if (v > best_dot) {
      best_dot = v;
      best = i;
    }
  }
  return V3{pts[3 * best], pts[3 * best + 1], pts[3 * best + 2]};
}

// a posed shape (query.h:43-72); `point` = PointSupport (query.h:74-107)
struct Posed {
  Shape s;
  PoseV pose;
  bool point;
  V3 p;
};
TRX_HD V3 supportWorld(const Posed &o, const V3 &d) {
  if (o.point) return o.p;
  return trxvm::pmulv(o.pose, supportLocal(o.s, qrot(qinv(o.pose.r), d)));
}

struct Vertex {
  V3 w;   // point of A (-) B
  V3 sa;  // the support point of A it came from
};
TRX_HD Vertex minkowski(const Posed &A, const Posed &B, const V3 &d) {
  Vertex v;
  v.sa = supportWorld(A, d);
  v.w = v.sa - supportWorld(B, -d);
  return v;
}

// ---- closest point of a simplex to the origin (Ericson, Real-Time Collision Detection 5.1: segment, triangle,
// tetrahedron by Voronoi regions).  On return the simplex holds only the supporting vertices, lam their
// barycentric weights, and the result is sum lam_i y_i.
TRX_HD V3 closestSegment(Vertex *y, double *lam, int &n) {
  const V3 a = y[0].w, b = y[1].w, ab = b - a;
  const double t = -dot(a, ab), den = dot(ab, ab);
  if (t <= 0.0 || den <= 0.0) {
    n = 1;
    lam[0] = 1.0;
    return a;
  }
  if (t >= den) {
    y[0] = y[1];
    n = 1;
    lam[0] = 1.0;
    return b;
  }
  const double s = t / den;
  lam[0] = 1.0 - s;
  lam[1] = s;
  return a + ab * s;
}
// i0, i1, i2: which vertices of y form the triangle; writes the kept vertices' indices and weights
TRX_HD V3 closestTriangleIdx(const Vertex *y, int i0, int i1, int i2, int *keep, double *lam, int &n) {
  const V3 a = y[i0].w, b = y[i1].w, c = y[i2].w;
  const V3 ab = b - a, ac = c - a;
  const double d1 = -dot(ab, a), d2 = -dot(ac, a);
  if (d1 <= 0.0 && d2 <= 0.0) {
    n = 1, keep[0] = i0, lam[0] = 1.0;
    return a;
  }
  const double d3 = -dot(ab, b), d4 = -dot(ac, b);
  if (d3 >= 0.0 && d4 <= d3) {
    n = 1, keep[0] = i1, lam[0] = 1.0;
    return b;
  }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
    const double v = d1 / (d1 - d3);
    n = 2, keep[0] = i0, keep[1] = i1, lam[0] = 1.0 - v, lam[1] = v;
    return a + ab * v;
  }
  const double d5 = -dot(ab, c), d6 = -dot(ac, c);
  if (d6 >= 0.0 && d5 <= d6) {
    n = 1, keep[0] = i2, lam[0] = 1.0;
    return c;
  }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
    const double w = d2 / (d2 - d6);
    n = 2, keep[0] = i0, keep[1] = i2, lam[0] = 1.0 - w, lam[1] = w;
    return a + ac * w;
  }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    n = 2, keep[0] = i1, keep[1] = i2, lam[0] = 1.0 - w, lam[1] = w;
    return b + (c - b) * w;
  }
  const double denom = 1.0 / (va + vb + vc);
  const double v = vb * denom, w = vc * denom;
  n = 3, keep[0] = i0, keep[1] = i1, keep[2] = i2, lam[0] = 1.0 - v - w, lam[1] = v, lam[2] = w;
  return a + ab * v + ac * w;
}
TRX_HD void keepVertices(Vertex *y, const int *keep, int n) {
  Vertex t[4];
  for (int i = 0; i < n; i++) t[i] = y[keep[i]];
  for (int i = 0; i < n; i++) y[i] = t[i];
}
// returns true when the origin lies inside the tetrahedron
TRX_HD bool closestTetrahedron(Vertex *y, double *lam, int &n, V3 &closest) {
  const int faces[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
  double best = 1e300;
  int best_keep[3] = {0, 0, 0}, best_n = 0;
  double best_lam[3] = {0, 0, 0};
  V3 best_point = V3{0, 0, 0};
  bool outside_any = false;
  for (int f = 0; f < 4; f++) {
    const V3 a = y[faces[f][0]].w, b = y[faces[f][1]].w, c = y[faces[f][2]].w, d = y[faces[f][3]].w;
    const V3 nrm = cross(b - a, c - a);
    const double sp = -dot(a, nrm);    // origin's side
    const double sd = dot(d - a, nrm); // the fourth vertex's side
    // a flat tetrahedron (sd ~ 0) has no inside: evaluate the face
    const double scale = sqrt(norm2(nrm)) * (sqrt(norm2(d - a)) + 1e-300);
    const bool flat = fabs(sd) <= 1e-14 * scale;
    if (flat || sp * sd < 0.0) {
      outside_any = true;
      int keep[3], kn;
      double l[3];
      const V3 q = closestTriangleIdx(y, faces[f][0], faces[f][1], faces[f][2], keep, l, kn);
      const double dist = norm2(q);
      if (dist < best) {
        best = dist;
        best_point = q;
        best_n = kn;
        for (int i = 0; i < kn; i++) best_keep[i] = keep[i], best_lam[i] = l[i];
      }
    }
  }
  if (!outside_any) return true;
  keepVertices(y, best_keep, best_n);
  n = best_n;
  for (int i = 0; i < n; i++) lam[i] = best_lam[i];
  closest = best_point;
  return false;
}

// ---- expanding polytope (van den Bergen, "Proximity queries and penetration depth computation on 3D game
// objects", GDC 2001) on A (-) B, started from a tetrahedron of support points
static constexpr int kEpaVertices = 64, kEpaFaces = 128, kEpaEdges = 48, kEpaCandidates = 32, kEpaIterations = 96;
static constexpr int kGjkIterations = 128;

struct EpaFace {
  V3 n;
  double h;
  unsigned char v[3];
  unsigned char alive;
};

TRX_HD bool epaMakeFace(const Vertex *vs, int a, int b, int c, EpaFace &f) {
  const V3 nrm = cross(vs[b].w - vs[a].w, vs[c].w - vs[a].w);
  const double len = sqrt(norm2(nrm));
  f.v[0] = (unsigned char)a, f.v[1] = (unsigned char)b, f.v[2] = (unsigned char)c, f.alive = 1;
  if (!(len > 0.0)) {
    f.n = V3{0, 0, 0};
    f.h = 1e300;  // never selected
    return false;
  }
  f.n = nrm * (1.0 / len);
  f.h = dot(f.n, vs[a].w);
  return true;
}

// four support points spanning a volume; returns false when A (-) B is flat
TRX_COL_FN bool epaSeed(const Posed &A, const Posed &B, Vertex *vs) {
  const V3 axes[3] = {V3{1, 0, 0}, V3{0, 1, 0}, V3{0, 0, 1}};
  // the two points farthest apart along an axis
  double best = -1.0;
  for (int k = 0; k < 3; k++) {
    const Vertex p = minkowski(A, B, axes[k]), q = minkowski(A, B, -axes[k]);
    const double d = norm2(p.w - q.w);
    if (d > best) best = d, vs[0] = p, vs[1] = q;
  }
  if (!(best > 0.0)) return false;
  const V3 e = vs[1].w - vs[0].w;
  // the point farthest from that line
  best = -1.0;
  for (int k = 0; k < 3; k++) {
    V3 dir = cross(e, axes[k]);
    if (!(norm2(dir) > 0.0)) continue;
    for (int sgn = 0; sgn < 2; sgn++) {
      const Vertex p = minkowski(A, B, sgn ? -dir : dir);
      const double d = norm2(cross(p.w - vs[0].w, e));
      if (d > best) best = d, vs[2] = p;
    }
  }
  if (!(best > 1e-30 * norm2(e))) return false;
  // and the point farthest from that plane
  const V3 nrm = cross(vs[1].w - vs[0].w, vs[2].w - vs[0].w);
  best = -1.0;
  for (int sgn = 0; sgn < 2; sgn++) {
    const Vertex p = minkowski(A, B, sgn ? -nrm : nrm);
    const double d = fabs(dot(p.w - vs[0].w, nrm));
    if (d > best) best = d, vs[3] = p;
  }
  return best > 1e-14 * norm2(nrm);
}

TRX_COL_FN void epaPenetration(const Posed &A, const Posed &B, Result &r) {
  Vertex vs[kEpaVertices];
  EpaFace fs[kEpaFaces];
  unsigned char edges[kEpaEdges][2];
  int nv = 4, nf = 0;
  r.status = 1;
  r.epa_iterations = 0;
  if (!epaSeed(A, B, vs)) {
    r.status = 2;
    return;
  }
  // orient the seed so that every face normal points away from the fourth vertex
  if (dot(cross(vs[1].w - vs[0].w, vs[2].w - vs[0].w), vs[3].w - vs[0].w) > 0.0) {
    const Vertex t = vs[1];
    vs[1] = vs[2];
    vs[2] = t;
  }
  epaMakeFace(vs, 0, 1, 2, fs[nf++]);
  epaMakeFace(vs, 0, 3, 1, fs[nf++]);
  epaMakeFace(vs, 0, 2, 3, fs[nf++]);
  epaMakeFace(vs, 1, 3, 2, fs[nf++]);
  int closest = 0;
  for (int it = 0; it < kEpaIterations; it++) {
    r.epa_iterations = it + 1;
    // the face nearest to the origin by signed distance (negative while the origin is still outside the polytope:
    // expanding the most violated face first encloses it)
    closest = -1;
    for (int i = 0; i < nf; i++)
      if (fs[i].alive && (closest < 0 || fs[i].h < fs[closest].h)) closest = i;
    const EpaFace f = fs[closest];
    const Vertex w = minkowski(A, B, f.n);
    const double reach = dot(w.w, f.n);
    if (reach - f.h <= 1e-12 + 1e-10 * fabs(f.h)) break;  // the face lies on the boundary of A (-) B
    if (nv >= kEpaVertices) break;
    bool duplicate = false;
    for (int i = 0; i < nv; i++) duplicate |= (norm2(vs[i].w - w.w) == 0.0);
    if (duplicate) break;
    // the faces the new point sees, grown from the closest face through shared edges: a connected patch has one
    // closed horizon whatever rounding decides for faces the point is coplanar with (cylinder caps, box sides)
    int cand[kEpaCandidates];
    int nc = 0;
    const double plane_eps = 1e-14 + 1e-12 * fabs(f.h);
    for (int i = 0; i < nf && nc < kEpaCandidates; i++)
      if (fs[i].alive && i != closest && dot(fs[i].n, w.w) - fs[i].h > plane_eps) cand[nc++] = i;
    fs[closest].alive = 2;
    for (bool grew = true; grew;) {
      grew = false;
      for (int c = 0; c < nc; c++) {
        EpaFace &g = fs[cand[c]];
        if (g.alive != 1) continue;
        bool adjacent = false;
        for (int j = 0; j < nf && !adjacent; j++) {
          if (fs[j].alive != 2) continue;
          for (int e = 0; e < 3; e++)
            for (int k = 0; k < 3; k++)
              adjacent |= (g.v[e] == fs[j].v[(k + 1) % 3] && g.v[(e + 1) % 3] == fs[j].v[k]);
        }
        if (adjacent) g.alive = 2, grew = true;
      }
    }
    int ne = 0;
    bool overflow = false;
    for (int i = 0; i < nf; i++) {
      if (fs[i].alive != 2) continue;
      fs[i].alive = 0;
      for (int e = 0; e < 3; e++) {
        const unsigned char a = fs[i].v[e], b = fs[i].v[(e + 1) % 3];
        int twin = -1;
        for (int k = 0; k < ne; k++)
          if (edges[k][0] == b && edges[k][1] == a) twin = k;
        if (twin >= 0) {
          edges[twin][0] = edges[ne - 1][0];
          edges[twin][1] = edges[ne - 1][1];
          ne--;
        } else if (ne < kEpaEdges) {
          edges[ne][0] = a, edges[ne][1] = b, ne++;
        } else {
          overflow = true;
        }
      }
    }
    if (overflow || ne == 0) {
      fs[closest].alive = 1;  // (keeps a valid answer; the polytope is abandoned)
      break;
    }
    vs[nv] = w;
    for (int k = 0; k < ne; k++) {
      // reuse a dead slot, else append
      int slot = -1;
      for (int i = 0; i < nf && slot < 0; i++)
        if (!fs[i].alive) slot = i;
      if (slot < 0) {
        if (nf >= kEpaFaces) {
          overflow = true;
          break;
        }
        slot = nf++;
      }
      epaMakeFace(vs, edges[k][0], edges[k][1], nv, fs[slot]);
    }
    nv++;
    if (overflow) break;
  }
  closest = -1;
  for (int i = 0; i < nf; i++)
    if (fs[i].alive && fs[i].h < 1e299 && (closest < 0 || fs[i].h < fs[closest].h)) closest = i;
  if (closest < 0) {
    r.status = 2;
    return;
  }
  const EpaFace f = fs[closest];
  // witness on A: the origin's projection onto the face in barycentric coordinates of its vertices
  const V3 p = f.n * f.h;
  const V3 a = vs[f.v[0]].w, b = vs[f.v[1]].w, c = vs[f.v[2]].w;
  const V3 v0 = b - a, v1 = c - a, v2 = p - a;
  const double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
  const double den = d00 * d11 - d01 * d01;
  double lv = 0.0, lw = 0.0;
  if (den > 0.0) {
    lv = (d11 * d20 - d01 * d21) / den;
    lw = (d00 * d21 - d01 * d20) / den;
  }
  const double lu = 1.0 - lv - lw;
  r.a = vs[f.v[0]].sa * lu + vs[f.v[1]].sa * lv + vs[f.v[2]].sa * lw;
  r.b = r.a - f.n * f.h;
  r.n = -f.n;
  r.d = -f.h;
}

// closest points of the cores of A and B (no margins)
TRX_COL_FN void coreQuery(const Posed &A, const Posed &B, Result &r) {
  Vertex y[4];
  double lam[4] = {1.0, 0.0, 0.0, 0.0};
  int n = 1;
  // guess = (1, 0, 0) as query.cpp:93
  y[0] = minkowski(A, B, V3{1.0, 0.0, 0.0});
  V3 v = y[0].w;
  double vv = norm2(v);
  bool inside = false;
  r.epa_iterations = 0;
  int it = 0;
  for (; it < kGjkIterations; it++) {
    if (!(vv > 1e-28)) {  // the cores touch or overlap
      inside = true;
      break;
    }
    const Vertex w = minkowski(A, B, -v);
    const double gap = vv - dot(v, w.w);  // >= |v|^2 - dist^2 ... an upper bound of what is left
    if (gap <= 1e-14 * vv || gap <= 1e-20) break;
    bool duplicate = false;
    for (int i = 0; i < n; i++) duplicate |= (norm2(y[i].w - w.w) == 0.0);
    if (duplicate) break;
    // (kept for the case that the enlarged simplex, a sliver, solves to a worse point in fp64)
    Vertex y_before[3];
    double lam_before[3];
    const int n_before = n;
    for (int i = 0; i < n; i++) y_before[i] = y[i], lam_before[i] = lam[i];
    y[n++] = w;
    V3 nv = v;
    if (n == 2) nv = closestSegment(y, lam, n);
    else if (n == 3) {
      int keep[3];
      double l[3];
      int kn;
      nv = closestTriangleIdx(y, 0, 1, 2, keep, l, kn);
      keepVertices(y, keep, kn);
      n = kn;
      for (int i = 0; i < n; i++) lam[i] = l[i];
    } else {
      if (closestTetrahedron(y, lam, n, nv)) {
        inside = true;
        break;
      }
    }
    // (the new simplex contains the old one's support, so |nv| <= |v| up to rounding)
    const double nvv = norm2(nv);
    if (!(nvv < vv)) {  // stalled in fp64: the previous point is the answer
      n = n_before;
      for (int i = 0; i < n; i++) y[i] = y_before[i], lam[i] = lam_before[i];
      break;
    }
    v = nv;
    vv = nvv;
  }
  r.gjk_iterations = it;
  if (inside) {
    epaPenetration(A, B, r);
    return;
  }
  V3 pa = V3{0, 0, 0};
  for (int i = 0; i < n; i++) pa = pa + y[i].sa * lam[i];
  // (when the loop ended without a sub-simplex solve lam is that of the last solve, whose point is v)
  r.status = 0;
  r.a = pa;
  r.b = pa - v;
  r.d = sqrt(vv);
  r.n = v * (1.0 / r.d);
}

// margins (sphere radii) on top of the core result: exact for point cores
TRX_HD void applyMargins(double ma, double mb, Result &r) {
  if (r.status == 2) {
    // no direction: the reference reports d = 1 and leaves the previous points (query.cpp:140-148); here a
    // fixed, finite answer: coincident witnesses, normal +x
    r.n = V3{1.0, 0.0, 0.0};
    r.d = -(ma + mb);
    r.b = r.a;
    return;
  }
  r.a = r.a - r.n * ma;
  r.b = r.b + r.n * mb;
  r.d -= ma + mb;
  if (r.d < 0.0 && r.status == 0) r.status = 1;
}

// doCollisionQuery over two posed shapes (ShapeCollisionPair::update, query.h:128-141: the normal is normalised
// once more there)
TRX_COL_FN void collideShapes(const double *heap, uint64_t pair_handle, const PoseV &pose_a, const PoseV &pose_b,
                              Result &r) {
  const double *pair = heap + pair_handle;
  Posed A, B;
  A.s = shapeAt(heap, (uint64_t)pair[PAIR_A]);
  B.s = shapeAt(heap, (uint64_t)pair[PAIR_B]);
  A.pose = pose_a, B.pose = pose_b;
  A.point = B.point = false;
  A.p = B.p = V3{0, 0, 0};
  if (A.s.kind == SHAPE_SPHERE && B.s.kind == SHAPE_SPHERE) {
    // two point cores: closed form
    const V3 ca = trxvm::pmulv(pose_a, A.s.center()), cb = trxvm::pmulv(pose_b, B.s.center());
    const V3 v = ca - cb;
    const double vv = norm2(v);
    r.gjk_iterations = r.epa_iterations = 0;
    r.a = ca, r.b = cb;
    if (vv > 1e-28) {
      r.status = 0, r.d = sqrt(vv), r.n = v * (1.0 / r.d);
    } else {
      r.status = 2;
    }
  } else {
    coreQuery(A, B, r);
  }
  applyMargins(A.s.margin(), B.s.margin(), r);
  const double nn = sqrt(norm2(r.n));
  r.n = r.n * (1.0 / fmax(1e-9, nn));  // normalized(), vector3.h
}

// CollisionShape::project: the shape at the identity pose against a point (query.h:111-121); polyhedra answer
// from their planes while the point is inside (shape.h:327-346)
TRX_COL_FN void projectPoint(const double *heap, uint64_t shape_handle, const V3 &point, V3 &normal, double &distance) {
  Posed A, B;
  A.s = shapeAt(heap, shape_handle);
  if (A.s.kind == SHAPE_POLYHEDRON) {
    bool inside = true;
    const double *pl = A.s.planes();
    for (int i = 0; i < A.s.n_planes; i++) {
      const double dist = pl[4 * i] * point.x + pl[4 * i + 1] * point.y + pl[4 * i + 2] * point.z + pl[4 * i + 3];
      if (dist > 0.0) inside = false;
      if (i == 0 || dist > distance) {
        distance = dist;
        normal = V3{-pl[4 * i], -pl[4 * i + 1], -pl[4 * i + 2]};
      }
    }
    if (inside) return;
  }
  A.pose.t = V3{0, 0, 0};
  A.pose.r = Q4{0, 0, 0, 1};
  A.point = false;
  A.p = V3{0, 0, 0};
  B.s = A.s;
  B.pose = A.pose;
  B.point = true;
  B.p = point;
  Result r;
  if (A.s.kind == SHAPE_SPHERE) {
    const V3 v = A.s.center() - point;
    const double vv = norm2(v);
    r.a = A.s.center(), r.b = point;
    if (vv > 1e-28) r.status = 0, r.d = sqrt(vv), r.n = v * (1.0 / r.d);
    else r.status = 2;
    r.gjk_iterations = r.epa_iterations = 0;
  } else {
    coreQuery(A, B, r);
  }
  applyMargins(A.s.margin(), 0.0, r);
  normal = r.n;
  distance = r.d;
}

}  // namespace trxcol
