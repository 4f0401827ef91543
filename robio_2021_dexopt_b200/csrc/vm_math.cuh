// vm_math.cuh -- fp64 value types and SE(3) arithmetic of the tape VM.
//
// Every formula keeps the operation order of the reference so that the only
// rounding differences against the CPU engine are FMA contractions:
//   quaternion * vector, quaternion * quaternion, normalized, angle-axis
//       tractor/include/tractor/geometry/quaternion.h:59-124,153-163
//   pose composition / inverse / residual
//       tractor/include/tractor/geometry/pose.h:36-41,55-64,138-155
//   manifold "plus" (orientation (+) rotation vector)
//       tractor/include/tractor/geometry/ops.h:925-956
#pragma once

#ifdef __CUDACC__
#define TRX_HD __host__ __device__ __forceinline__
#else
#define TRX_HD inline
#endif

#include <math.h>

namespace trxvm {

struct V3 {
  double x, y, z;
};
struct Q4 {
  double x, y, z, w;
};
struct PoseV {
  V3 t;
  Q4 r;
};
struct TwistV {
  V3 t;
  V3 r;
};

TRX_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
TRX_HD V3 operator+(const V3 &a, const V3 &b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
TRX_HD V3 operator-(const V3 &a, const V3 &b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
TRX_HD V3 operator-(const V3 &a) { return V3{-a.x, -a.y, -a.z}; }
TRX_HD V3 operator*(const V3 &a, double b) { return V3{a.x * b, a.y * b, a.z * b}; }
TRX_HD V3 operator*(double a, const V3 &b) { return V3{a * b.x, a * b.y, a * b.z}; }
TRX_HD double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
TRX_HD V3 cross(const V3 &a, const V3 &b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

TRX_HD Q4 qinv(const Q4 &q) { return Q4{-q.x, -q.y, -q.z, q.w}; }

// quaternion.h:71-104
TRX_HD V3 qrot(const Q4 &q, const V3 &v) {
  double t_x = q.y * v.z - q.z * v.y;
  double t_y = q.z * v.x - q.x * v.z;
  double t_z = q.x * v.y - q.y * v.x;
  double r_x = q.w * t_x + q.y * t_z - q.z * t_y;
  double r_y = q.w * t_y + q.z * t_x - q.x * t_z;
  double r_z = q.w * t_z + q.x * t_y - q.y * t_x;
  r_x += r_x;
  r_y += r_y;
  r_z += r_z;
  r_x += v.x;
  r_y += v.y;
  r_z += v.z;
  return V3{r_x, r_y, r_z};
}

// quaternion.h:106-124
TRX_HD Q4 qmul(const Q4 &p, const Q4 &q) {
  Q4 r;
  r.x = (p.w * q.x + p.x * q.w) + (p.y * q.z - p.z * q.y);
  r.y = (p.w * q.y - p.x * q.z) + (p.y * q.w + p.z * q.x);
  r.z = (p.w * q.z + p.x * q.y) - (p.y * q.x - p.z * q.w);
  r.w = (p.w * q.w - p.x * q.x) - (p.y * q.y + p.z * q.z);
  return r;
}

// quaternion.h:59-69
TRX_HD Q4 qnormalized(const Q4 &q) {
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  double f = 1.0 / fmax(1e-9, n);
  return Q4{q.x * f, q.y * f, q.z * f, q.w * f};
}

// quaternion.h:153-163
TRX_HD Q4 qangleaxis(double angle, const V3 &axis) {
  double s, c;
#ifdef __CUDA_ARCH__
  sincos(angle * 0.5, &s, &c);
#else
  s = sin(angle * 0.5);
  c = cos(angle * 0.5);
#endif
  return Q4{axis.x * s, axis.y * s, axis.z * s, c};
}

// geometry/ops.h:951-956
TRX_HD Q4 qplus(const Q4 &a, const V3 &b) {
  return qnormalized(qmul(qnormalized(Q4{b.x * 0.5, b.y * 0.5, b.z * 0.5, 1.0}), a));
}

// quaternion.h:140-151 (sign chosen per lane from w)
TRX_HD V3 qresidual(const Q4 &a) {
  double f = (a.w < 0) ? -2.0 : 2.0;
  return V3{a.x * f, a.y * f, a.z * f};
}

// pose.h:55-64
TRX_HD PoseV pmul(const PoseV &a, const PoseV &b) {
  PoseV x;
  x.t = a.t + qrot(a.r, b.t);
  x.r = qmul(a.r, b.r);
  return x;
}
TRX_HD V3 pmulv(const PoseV &a, const V3 &b) { return a.t + qrot(a.r, b); }

}  // namespace trxvm
