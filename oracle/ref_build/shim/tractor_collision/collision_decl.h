// Declarations that let lines 206-375 of /root/reference/dexopt/src/ops.h (collision_project_2, collision_project,
// collision_axes: operator bodies, derivative rules, registration) compile in this image.  The real classes
// (tractor/include/tractor/collision/shape.h:14-50, collision/query.h:123-152) need Eigen and Bullet, neither of which
// is here; these stand-ins have the members the operators touch and throw when a query is asked for.  Purpose: the
// reference's operator registry then lists the collision operators, which pins their names and argument layouts
// (tests/golden/ref_registry.txt); their results stay unpinned (no Bullet).
#pragma once
#include <stdexcept>
#include <tractor/geometry/pose.h>
#include <tractor/geometry/twist.h>
#include <tractor/geometry/vector3.h>
namespace tractor {
template <class Scalar> class CollisionShape {
public:
  virtual ~CollisionShape() {}
  virtual void project(const Vector3<Scalar> &, Vector3<Scalar> &, Scalar &) {
    throw std::runtime_error("collision queries need Bullet, which this image does not have");
  }
};
template <class Scalar> class ShapeCollisionPair {
  Vector3<Scalar> _point_a, _point_b, _normal;

public:
  void update(const Pose<Scalar> &, const Pose<Scalar> &) {
    throw std::runtime_error("collision queries need Bullet, which this image does not have");
  }
  const Vector3<Scalar> &pointA() const { return _point_a; }
  const Vector3<Scalar> &pointB() const { return _point_b; }
  const Vector3<Scalar> &normal() const { return _normal; }
};
}  // namespace tractor
