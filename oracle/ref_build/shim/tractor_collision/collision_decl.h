// Stand-ins that let lines 206-375 of /root/reference/dexopt/src/ops.h (collision_project_2, collision_project,
// collision_axes: operator bodies, derivative rules, registration) compile AND run in this image.
//
// The real classes (tractor/include/tractor/collision/shape.h:14-50, collision/query.h:123-152) need Eigen and end in
// Bullet's btGjkEpaSolver3 (src/collision/query.cpp:93-105); neither library is here.  These stand-ins have exactly the
// members the operators touch.  The one call the reference hands to Bullet -- "closest points of two posed convex
// shapes" -- goes through a hook that the test tool fills in (ref_tool.cpp: the geometric query under test).  So on a
// tape recorded through the reference API, the REFERENCE's own code decides everything around that call: operator
// bodies (local = pose^-1 * point, normalisation), derivative rules, buildGradients, SimpleEngine.  What stays
// unpinned is the query's own arithmetic against Bullet's.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <tractor/geometry/pose.h>
#include <tractor/geometry/twist.h>
#include <tractor/geometry/vector3.h>
namespace tractor {
struct CollisionQueryHook {
  // shape handle, point xyz -> normal xyz, distance            (CollisionShape::project, query.h:111-121)
  void (*project)(uint64_t shape, const double *point, double *normal, double *distance) = nullptr;
  // pair handle, poses (translation xyz + quaternion xyzw) -> point a, point b, normal (doCollisionQuery's result)
  void (*pair)(uint64_t pair, const double *pose_a, const double *pose_b, double *a, double *b, double *n) = nullptr;
};
inline CollisionQueryHook &collisionQueryHook() {
  static CollisionQueryHook hook;
  return hook;
}
template <class Scalar> class CollisionShape {
public:
  uint64_t handle = 0;
  virtual ~CollisionShape() {}
  virtual void project(const Vector3<Scalar> &point, Vector3<Scalar> &normal, Scalar &distance) {
    if (!collisionQueryHook().project) throw std::runtime_error("collision queries need Bullet, which this image does not have");
    const double p[3] = {double(point.x()), double(point.y()), double(point.z())};
    double n[3], d;
    collisionQueryHook().project(handle, p, n, &d);
    normal = Vector3<Scalar>(Scalar(n[0]), Scalar(n[1]), Scalar(n[2]));
    distance = Scalar(d);
  }
};
template <class Scalar> class ShapeCollisionPair {
  Vector3<Scalar> _point_a, _point_b, _normal;

public:
  uint64_t handle = 0;
  void update(const Pose<Scalar> &pose_a, const Pose<Scalar> &pose_b) {
    if (!collisionQueryHook().pair) throw std::runtime_error("collision queries need Bullet, which this image does not have");
    auto flat = [](const Pose<Scalar> &p, double *o) {
      o[0] = double(p.translation().x()), o[1] = double(p.translation().y()), o[2] = double(p.translation().z());
      o[3] = double(p.orientation().x()), o[4] = double(p.orientation().y()), o[5] = double(p.orientation().z());
      o[6] = double(p.orientation().w());
    };
    double pa[7], pb[7], a[3], b[3], n[3];
    flat(pose_a, pa);
    flat(pose_b, pb);
    collisionQueryHook().pair(handle, pa, pb, a, b, n);
    _point_a = Vector3<Scalar>(Scalar(a[0]), Scalar(a[1]), Scalar(a[2]));
    _point_b = Vector3<Scalar>(Scalar(b[0]), Scalar(b[1]), Scalar(b[2]));
    _normal = normalized(Vector3<Scalar>(Scalar(n[0]), Scalar(n[1]), Scalar(n[2])));  // as query.h:137-138
  }
  const Vector3<Scalar> &pointA() const { return _point_a; }
  const Vector3<Scalar> &pointB() const { return _point_b; }
  const Vector3<Scalar> &normal() const { return _normal; }
};
}  // namespace tractor
