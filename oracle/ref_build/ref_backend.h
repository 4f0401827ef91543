// ref_backend.h -- assembly backend that records through the REFERENCE's own
// Var<> / operator API (tractor/include/tractor/core/var.h:12-81,
// core/operator.h:295-316), so every arithmetic step of the oracle runs inside
// reference code.  Used with csrc/assembly/dexsyn_assemble.h.
//
// TEST INFRASTRUCTURE ONLY (compiled into oracle/_ref/ref_tool).
#pragma once
#include <deque>

#include "ref_common.h"

namespace refc {

template <class Lane> struct RefBackend {
  // collision_axes / collision_project_2 are the reference's operators (dexopt/src/ops.h:208-375) over the stand-in
  // shape classes of shim/tractor_collision/collision_decl.h; the uint64 the assembly passes is the address of such
  // an object, as in the reference (dexlearn.h:85-93,329)
  static constexpr bool kHasCollisionOps = true;
  typedef Lane LaneValue;  // tractor::Batch8d
  typedef tractor::Var<Lane> S;
  typedef tractor::Var<double> W;
  typedef tractor::Var<tractor::Vector3<Lane>> V3;
  typedef tractor::Var<tractor::Quaternion<Lane>> Q;
  typedef tractor::Var<tractor::Pose<Lane>> P;
  typedef tractor::Var<tractor::Matrix3<Lane>> M3;

  // stable storage for ports (an input/parameter Var must keep its host address,
  // recorder.cpp:498-539)
  std::deque<W, tractor::AlignedStdAlloc<W>> weights;
  std::deque<S, tractor::AlignedStdAlloc<S>> params;

  template <class T> static T &nc(const T &v) { return const_cast<T &>(v); }

  // ---- constants
  S cs(double v) { return S(Lane(v)); }
  W cw(double v) { return W(v); }
  V3 cv(double x, double y, double z) { return V3(tractor::Vector3<Lane>(Lane(x), Lane(y), Lane(z))); }
  Q cq(double x, double y, double z, double w) {
    return Q(tractor::Quaternion<Lane>(Lane(x), Lane(y), Lane(z), Lane(w)));
  }
  P cp(const double *o) {
    return P(tractor::Pose<Lane>(tractor::Vector3<Lane>(Lane(o[0]), Lane(o[1]), Lane(o[2])),
                                 tractor::Quaternion<Lane>(Lane(o[3]), Lane(o[4]), Lane(o[5]), Lane(o[6]))));
  }
  M3 cm(const double *m) {
    tractor::Matrix3<Lane> r;
    for (int row = 0; row < 3; row++)
      for (int col = 0; col < 3; col++) r(row, col) = Lane(m[row * 3 + col]);
    return M3(r);
  }

  // ---- ports
  W weight(double init) {  // neural.h:146-157: Var<double>, registered with variable()
    weights.emplace_back(init);
    tractor::variable(weights.back());
    return weights.back();
  }
  S param() {  // dexenv.h:47-55: per-lane value fed through Executable::parameterize
    params.emplace_back(Lane(0));
    tractor::parameter(params.back());
    return params.back();
  }
  void goal(const S &v) { tractor::goal(v); }
  void goalV(const V3 &v) { tractor::goal(v); }
  void goalW(const W &v) { tractor::goal(v); }

  void frameBegin(int) {}
  void frameEnd(int) {}

  // ---- scalar ops (core/ops.h:18-152, dexopt ops.h:99-116)
  S add(const S &a, const S &b) { return tractor::add(a, b); }
  S sub(const S &a, const S &b) { return tractor::sub(a, b); }
  S mul(const S &a, const S &b) { return tractor::mul(a, b); }
  S div(const S &a, const S &b) { return tractor::div(a, b); }
  S neg(const S &a) { return tractor::minus(a); }
  S tanh(const S &a) { return tractor::tanh(a); }
  // dexopt/src/ops.h:99 also declares a plain template relu(const T&); a non-const lvalue
  // selects the recording overload, as the reference's own call sites do (dexlearn.h:150-153)
  S relu(const S &a) { return tractor::relu(nc(a)); }
  S exp(const S &a) { return tractor::exp(a); }
  S log(const S &a) { return tractor::log(a); }
  W wmul(const W &a, const W &b) { return tractor::mul(a, b); }

  // ---- vectors (geometry/ops.h:84-280)
  V3 vadd(const V3 &a, const V3 &b) { return tractor::add(a, b); }
  V3 vsub(const V3 &a, const V3 &b) { return tractor::sub(a, b); }
  V3 vneg(const V3 &a) { return tractor::minus(a); }
  V3 vmuls(const V3 &a, const S &b) { return tractor::mul(a, b); }
  S dot(const V3 &a, const V3 &b) { return tractor::dot(a, b); }
  V3 cross(const V3 &a, const V3 &b) { return tractor::cross(a, b); }
  V3 pack(const S &x, const S &y, const S &z) {
    V3 r;
    tractor::fg_vec3_pack(x, y, z, r);
    return r;
  }
  void unpack(const V3 &v, S &x, S &y, S &z) { tractor::fg_vec3_unpack(v, x, y, z); }

  // ---- orientations (geometry/ops.h:313-443,951-970)
  V3 qmulv(const Q &q, const V3 &v) { return tractor::mul(q, v); }
  Q qmul(const Q &a, const Q &b) { return tractor::mul(a, b); }
  Q qinv(const Q &a) { return tractor::fg_quat_inverse(a); }
  Q qaxis(const S &angle, const V3 &axis) { return tractor::fg_angle_axis_quat(angle, axis); }
  Q qaddv(const Q &q, const V3 &v) { return tractor::add(q, v); }
  V3 qresidual(const Q &q) { return tractor::fg_quat_residual(q); }

  // ---- poses (geometry/ops.h:520-786)
  P pmul(const P &a, const P &b) { return tractor::mul(a, b); }
  V3 pmulv(const P &a, const V3 &v) { return tractor::mul(a, v); }
  P prevolute(const P &parent, const S &angle, const V3 &axis) {
    return tractor::fg_pose_angle_axis_pose(parent, angle, axis);
  }
  P ptrans(const V3 &t) { return tractor::fg_translation_pose(t); }
  P porient(const Q &q) { return tractor::fg_orientation_pose(q); }
  V3 ptranslation(const P &p) { return tractor::fg_pose_translation(p); }
  Q porientation(const P &p) { return tractor::fg_pose_orientation(p); }
  V3 mmulv(const M3 &m, const V3 &v) { return tractor::mul(m, v); }

  // ---- collision operators (dexopt/src/ops.h:208-252,288-375), called as dexlearn.h:93-95,329-330 call them
  void collisionProject(const V3 &point, uint64_t shape, V3 &normal, S &distance) {
    tractor::collision_project_2(nc(point), shape, normal, distance);
  }
  void collisionAxes(const P &pose_a, const P &pose_b, uint64_t pair, V3 &point_a, V3 &point_b, V3 &axis, V3 &local_a,
                     V3 &local_b) {
    tractor::collision_axes(nc(pose_a), nc(pose_b), pair, point_a, point_b, axis, local_a, local_b);
  }

  // ---- dense layer: tensor_mul_vec_mat (dexopt/src/tensor.h:134-172) + bias
  // broadcast/add (dexopt/src/neural.h:176-181), weights[row * n_out + col]
  std::vector<S> dense(const std::vector<S> &a, const std::vector<W> &wts, const std::vector<W> &bias,
                       int n_in, int n_out) {
    std::vector<S> r;
    for (int col = 0; col < n_out; col++) {
      S acc;  // default Var<Batch>: a zero constant, as Tensor::resize does
      int row = 0;
      for (; row + 3 < n_in; row += 4) {
        S w0, w1, w2, w3;
        tractor::batch(wts[(row + 0) * n_out + col], w0);
        tractor::batch(wts[(row + 1) * n_out + col], w1);
        tractor::batch(wts[(row + 2) * n_out + col], w2);
        tractor::batch(wts[(row + 3) * n_out + col], w3);
        acc = tractor::dot4add(nc(a[row + 0]), nc(a[row + 1]), nc(a[row + 2]), nc(a[row + 3]), w0, w1, w2, w3, acc);
      }
      for (; row < n_in; row++) {
        S w;
        tractor::batch(wts[row * n_out + col], w);
        acc = tractor::madd(nc(a[row]), w, acc);
      }
      r.push_back(acc);
    }
    for (int i = 0; i < n_out; i++) {
      S bv;
      tractor::batch(bias[i], bv);
      r[i] = tractor::add(r[i], bv);
    }
    return r;
  }
};

}  // namespace refc
