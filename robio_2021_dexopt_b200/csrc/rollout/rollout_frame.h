// rollout_frame.h -- one frame of dexopt's rollout objective, forward and adjoint, hand-written.
//
// What the reference executes per frame as ~1 500 interpreted tape instructions (after dense-layer
// fusion; ~12 000 before) is written out here as a fixed sequence of *phases*.  Inside a phase every
// item (a joint, a policy output, one side of a contact) is independent of the others and touches only
// its own storage, so whoever runs the phase may deal the items of a rollout to lanes in any way
// (`lane` = first item, stride 32) and only has to synchronise between phases.  The kernel
// (trx_rollout.cu) puts the same item of the 8 rollouts of a CTA into neighbouring lanes; the forward
// kinematics and its adjoint are warp scans instead (fkRound*: one warp per rollout, a lane per joint),
// the body step is one lane per rollout (phaseBody / bphaseBody), and the policy layers are the one
// place where the rollouts of a CTA meet (DMMA tiles over rollouts, trx_rollout.cu).
//
// Forward follows csrc/assembly/dexsyn_assemble.h statement by statement, which in turn restates
//   simulator step        dexopt/src/physics5.h:106-121,174-193,334-391,572-589
//   forward kinematics    tractor/include/tractor/robot/robotmodel.h:128-148, robot/jointtypes.h:143-153,296-301
//   joint limits          dexopt/src/dexlearn.h:139-157
//   observation, control  dexopt/src/dexenv_grasp.h:50-121,186-220
//   contacts, penalties   dexopt/src/dexlearn.h:159-217,235-391
//   terminal goals        dexopt/src/goals.h:71-89,173-191
// The adjoint mirrors the same statements in reverse with the reference's derivative rules AS WRITTEN
// (tractor/include/tractor/geometry/ops.h; e.g. fg_quat_inverse negates the tangent, fg_quat_residual and
// add_fg_quat_vec3 pass it through, relu tests a >= 0), not with the true derivatives: the rule each
// statement uses is named next to it.  Nothing is saved per operator: the adjoint recomputes the
// frame from the per-frame checkpoint (joint positions, commands, body state, tracked link poses).
//
// One deliberate re-association: the reference composes parent * origin and then rotates about the
// joint axis (fg_pose_angle_axis_pose); here origin * R(axis, q) is formed per joint first (all joints in
// parallel) and the chain is one pose product per joint.  Same value up to rounding, same derivative.
//
// Everything is __host__ __device__: tests/emu/rollout_emu.cpp runs the same phases lane by lane on
// the CPU to check them against the reference goldens without a GPU (test infrastructure only).
#pragma once
#include <cstdint>

#include "../vm_math.cuh"
#include "rollout_plan.h"

namespace rollout {
using trxvm::cross;
using trxvm::dot;
using trxvm::pmulv;
using trxvm::PoseV;
using trxvm::Q4;
using trxvm::qangleaxis;
using trxvm::qinv;
using trxvm::qmul;
using trxvm::qplus;
using trxvm::qresidual;
using trxvm::qrot;
using trxvm::TwistV;
using trxvm::V3;

#define RL_HD TRX_HD

struct Ctx {
  const int32_t *H;  // plan header (kernel-parameter constant bank on the device; == I on the host)
  const int32_t *I;  // plan tables (shared memory on the device)
  const double *D;
  double *s;         // this rollout's working region
  double *x, *h, *a, *out;    // this rollout's rows of the CTA's activation matrices
  double *gx, *gh, *gout;     // and of their adjoints
  double *r_out;     // residual vector of this rollout, element stride r_stride (null: not stored)
  const double *seed;  // adjoint seeds by residual row (null: the residuals themselves, g = J^T r)
  long r_stride;
  int frames;
};

// ---- small load/store helpers
RL_HD V3 ld3(const double *p) { return V3{p[0], p[1], p[2]}; }
RL_HD Q4 ld4(const double *p) { return Q4{p[0], p[1], p[2], p[3]}; }
RL_HD PoseV ldp(const double *p) { return PoseV{{p[0], p[1], p[2]}, {p[3], p[4], p[5], p[6]}}; }
RL_HD TwistV ldt(const double *p) { return TwistV{{p[0], p[1], p[2]}, {p[3], p[4], p[5]}}; }
RL_HD void st3(double *p, const V3 &v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }
RL_HD void st4(double *p, const Q4 &q) { p[0] = q.x, p[1] = q.y, p[2] = q.z, p[3] = q.w; }
RL_HD void stp(double *p, const PoseV &v) { st3(p, v.t), st4(p + 3, v.r); }
RL_HD void stt(double *p, const TwistV &v) { st3(p, v.t), st3(p + 3, v.r); }
RL_HD void add3(double *p, const V3 &v) { p[0] += v.x, p[1] += v.y, p[2] += v.z; }
RL_HD void addt(double *p, const TwistV &v) { add3(p, v.t), add3(p + 3, v.r); }
RL_HD double relu(double a) { return fmax(0.0, a); }

RL_HD const double *dconst(const Ctx &c) { return c.D + c.H[H_D_CONST]; }

// residual row: value into the loss (and the residual vector when asked for)
RL_HD void emit(const Ctx &c, long row, double v, double &loss) {
  loss += v * v;
  if (c.r_out) c.r_out[row * c.r_stride] = v;
}
// adjoint seed of a residual row: the residual itself (g = J^T r, solvers/gd.h:60-63) or a given vector (J^T v)
RL_HD double seedOf(const Ctx &c, long row, double v) { return c.seed ? c.seed[row * c.r_stride] : v; }

// GeometryFast::inverse(Pose) (geometry/fast.h:143-147): ri = inverse(orientation), ti = ri * (-translation);
// pmul(ptrans(ti), porient(ri)) adds zero and multiplies by the identity quaternion: exactly (ti, ri)
RL_HD PoseV pinv(const PoseV &p) {
  Q4 ri = qinv(p.r);
  return PoseV{qrot(ri, -p.t), ri};
}
// adjoint of pinv, as the tape composes it: fg_pose_translation / fg_pose_orientation, fg_quat_inverse
// (tangent negated), minus_fg_vec3, mul_fg_quat_vec3, fg_translation_pose / fg_orientation_pose, mul_fg_pose
RL_HD TwistV pinvReverse(const PoseV &p, const PoseV &inv, const TwistV &g_inv) {
  V3 g_ti = g_inv.t;                       // reverse_mul_fg_pose da.t, reverse_fg_translation_pose
  V3 g_ri = g_inv.r;                       // db.r = qrot(qinv(identity), dx.r), reverse_fg_orientation_pose
  g_ri = g_ri + cross(inv.t, g_ti);        // reverse_mul_fg_quat_vec3: da = cross(qrot(va, vb), dx)
  V3 g_negt = qrot(p.r, g_ti);             //                           db = qrot(qinv(va), dx), qinv(ri) = p.r
  return TwistV{-g_negt, -g_ri};           // reverse_minus_fg_vec3; reverse_fg_quat_inverse
}

// ================================================================================== forward phases

// start of a forward frame: the contact impulses of the previous frame reach the momenta (applyForce,
// physics5.h:114-121,672-675, in end-effector order), the link poses of the previous frame become "previous"
// (physics5.h:577), and the carried state {q, commands, body, previous poses} is checkpointed -- one contiguous copy
RL_HD void phaseBeginFrame(const Ctx &c, int lane, int frame, double *ckpt) {
  const int32_t *I = c.I, *H = c.H;
  const int nq = H[H_NQ], nt = H[H_NTRACK];
  double *s = c.s;
  double *carried = s + H[R_Q];
  const int n_state = 2 * nq + B_COUNT;
  for (int e = lane; e < n_state; e += 32) {
    double v = carried[e];
    const int m = e - (2 * nq + B_LM);  // 0..5: linear, angular momentum components
    if (m >= 0 && frame > 0) {
      for (int ee = 0; ee < H[H_NEE]; ee++) v += s[H[R_EE] + ee * EE_STRIDE + EE_IMP + m];
      carried[e] = v;
    }
    ckpt[e] = v;
  }
  for (int e = lane; e < nt * 7; e += 32) {
    const int t = e / 7, k = I[H[H_I_TRACK] + t];
    const double v = s[H[R_POSE] + k * 7 + (e - t * 7)];
    carried[n_state + e] = v;
    ckpt[n_state + e] = v;
  }
}

// start of an adjoint frame: restore the carried state the frame started from (values prefetched into
// registers by the caller: pre[i] = ckpt[lane + 32 i])
template <int N> RL_HD void phaseLoadCheckpoint(const Ctx &c, int lane, const double (&pre)[N]) {
  double *carried = c.s + c.H[R_Q];
  const int n = c.H[H_CKPT];
#pragma unroll
  for (int i = 0; i < N; i++)
    if (lane + 32 * i < n) carried[lane + 32 * i] = pre[i];
}

// intermediate values of the body step that both sweeps need
struct BodyStep {
  V3 pos, lm, am;  // state at frame start
  Q4 ori;
  V3 lm2, am2, lam, w, av, rot, rc, pos2, am3;
  Q4 ori2;
  PoseV A, Bp, C, Dp;
  V3 cg;
};

// PhysicsSimulator::step for the one dynamic body (physics5.h:572-589 with the contact detection empty)
RL_HD void bodyForward(const double *dc, const double *body, BodyStep &b) {
  b.pos = ld3(body + B_POS), b.ori = ld4(body + B_ORI), b.lm = ld3(body + B_LM), b.am = ld3(body + B_AM);
  const V3 center = ld3(dc + DC_CENTER);
  const double dt = dc[DC_DT];
  b.cg = b.pos + qrot(b.ori, center);                  // _updateBodyPoses
  V3 lm1 = b.lm + ld3(dc + DC_GRAV);                    // _applyGravity (gravity * (mass * dt) is a folded constant)
  b.lm2 = lm1 * dc[DC_DAMP_LIN];                        // _applyDamping
  b.am2 = b.am * dc[DC_DAMP_ANG];
  b.lam = qrot(qinv(b.ori), b.am2);                     // _integrate: local angular momentum
  V3 lv = b.lm2 * dc[DC_INV_MASS];
  V3 pos1 = b.pos + lv * dt;
  {
    const double *m = dc + DC_IINV;                     // mul_fg_mat3_vec3: rows accumulate left to right from zero
    double x[3] = {0, 0, 0}, v[3] = {b.lam.x, b.lam.y, b.lam.z};
    for (int row = 0; row < 3; row++)
      for (int col = 0; col < 3; col++) x[row] += m[row * 3 + col] * v[col];
    b.w = V3{x[0], x[1], x[2]};
  }
  b.av = qrot(b.ori, b.w);
  b.rot = b.av * dt;
  b.rc = qrot(b.ori, center);
  b.pos2 = pos1 - cross(b.rot, b.rc);
  b.ori2 = qplus(b.ori, b.rot);                         // add_fg_quat_vec3
  b.am3 = qrot(b.ori2, b.lam);
  // _updateRobotState (physics5.h:360-371): inverse_anchor * T(position) * R(orientation)
  const PoseV ia = ldp(dc + DC_ANCHOR);
  b.A = trxvm::pmul(ia, PoseV{b.pos2, {0, 0, 0, 1}});
  b.Bp = trxvm::pmul(b.A, PoseV{{0, 0, 0}, b.ori2});
  // floating joint (jointtypes.h:296-301): origin * T(float_position) * R(float_orientation)
  const PoseV org = ldp(dc + DC_OBJ_ORIGIN);
  b.C = trxvm::pmul(org, PoseV{b.Bp.t, {0, 0, 0, 1}});
  b.Dp = trxvm::pmul(b.C, PoseV{{0, 0, 0}, b.Bp.r});
}

// simulator step, joints: the velocity controllers (physics5.h:373-391) and the local joint rotations
// origin.r * R(axis, q)
RL_HD void phaseSim(const Ctx &c, int lane) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  const int nq = H[H_NQ];
  for (int ci = lane; ci < nq; ci += 32) {
    double q = s[H[R_Q] + ci] + s[H[R_CMD] + ci] * dc[DC_DT];
    s[H[R_Q] + ci] = q;
    const int k = I[H[H_I_CTRL] + ci * CI_STRIDE + CI_JOINT];
    const double *jd = c.D + H[H_D_JOINT] + k * JD_STRIDE;
    st4(s + H[R_LQ] + k * 4, qmul(ld4(jd + JD_ORIGIN + 3), qangleaxis(q, ld3(jd + JD_AXIS))));
  }
}
// what is left of the simulator step when the adjoint sweep takes a frame's values from its checkpoint instead of
// recomputing them: the joints advance, and the body state the frame started from is the carried one
RL_HD void phaseAdvance(const Ctx &c, int lane) {
  const int32_t *H = c.H;
  const double dt = dconst(c)[DC_DT];
  double *s = c.s;
  for (int ci = lane; ci < H[H_NQ]; ci += 32) s[H[R_Q] + ci] += s[H[R_CMD] + ci] * dt;
  if (lane == 31)
    for (int e = 0; e < B_COUNT; e++) s[H[R_BODY0] + e] = s[H[R_BODY] + e];
}
// simulator step, body (one lane per rollout): integrates the dynamic body and poses the object link
RL_HD void phaseBody(const Ctx &c) {
  const int32_t *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  {
    BodyStep b;
    double *body = s + H[R_BODY];
    for (int e = 0; e < B_COUNT; e++) s[H[R_BODY0] + e] = body[e];  // the adjoint restarts from here
    bodyForward(dc, body, b);
    st3(s + H[R_CG], b.cg);
    st3(body + B_POS, b.pos2);
    st4(body + B_ORI, b.ori2);
    st3(body + B_LM, b.lm2);
    st3(body + B_AM, b.am3);
    stp(s + H[R_POSE] + H[H_OBJ] * 7, b.Dp);
  }
}

// ---- forward kinematics (pose = parent * (origin.t, local rotation), robotmodel.h:128-148; the reference rotates
// about the joint axis after composing parent * origin, see the header comment) as warp scans (rollout_plan.h, RD_*): one warp owns one rollout, a lane per joint
// of a *round*; the chain products and sums inside a segment are prefix scans over the lanes.  Written over N lanes
// at once with the lane exchange as a policy: N = 1 and warp shuffles on the device (trx_rollout.cu), N = 32 and
// array shifts in the host emulation (tests/emu/rollout_emu.cpp) -- same statements, same order.
//   W::up(x, d, o):   o[lane] = x[lane - d]   (lanes below d keep their own value)
//   W::down(x, d, o): o[lane] = x[lane + d]
struct RoundLane {
  int k, j, len, par;
};
RL_HD RoundLane roundLane(const Ctx &c, int round, int lane) {
  const int32_t *rd = c.I + c.H[H_I_ROUND] + (round * 32 + lane) * RD_STRIDE;
  return RoundLane{rd[RD_K], rd[RD_J], rd[RD_LEN], rd[RD_PAR]};
}

template <int N, class W> RL_HD void fkRoundForward(const Ctx &c, int round, int lane0) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  const int steps = I[H[H_I_ROUNDSTEPS] + round];
  RoundLane rl[N];
  Q4 q[N], o4[N], r[N];
  V3 a[N], o3[N];
  PoseV pp[N];
  for (int i = 0; i < N; i++) {
    rl[i] = roundLane(c, round, lane0 + i);
    const bool on = rl[i].k >= 0;
    q[i] = on ? ld4(s + H[R_LQ] + rl[i].k * 4) : Q4{0, 0, 0, 1};
    pp[i] = on ? ldp(s + H[R_POSE] + rl[i].par * 7) : PoseV{{0, 0, 0}, {0, 0, 0, 1}};
  }
  // q_j = l_0 * l_1 * ... * l_j  (local rotations of the segment up to this joint)
  for (int st = 0, d = 1; st < steps; st++, d <<= 1) {
    W::up(q, d, o4);
    for (int i = 0; i < N; i++)
      if (rl[i].j >= d) q[i] = qmul(o4[i], q[i]);
  }
  for (int i = 0; i < N; i++) r[i] = qmul(pp[i].r, q[i]);
  // arm of joint j: rotation of its parent (the lane below, or the segment's parent) applied to its origin
  W::up(r, 1, o4);
  for (int i = 0; i < N; i++) {
    const Q4 rpar = rl[i].j == 0 ? pp[i].r : o4[i];
    a[i] = rl[i].k >= 0 ? qrot(rpar, ld3(c.D + H[H_D_JOINT] + rl[i].k * JD_STRIDE + JD_ORIGIN)) : V3{0, 0, 0};
  }
  for (int st = 0, d = 1; st < steps; st++, d <<= 1) {
    W::up(a, d, o3);
    for (int i = 0; i < N; i++)
      if (rl[i].j >= d) a[i] = o3[i] + a[i];
  }
  for (int i = 0; i < N; i++)
    if (rl[i].k >= 0) stp(s + H[R_POSE] + rl[i].k * 7, PoseV{pp[i].t + a[i], r[i]});
}

// joint-limit penalties (dexlearn.h:139-157) and the policy input (dexenv_grasp.h:50-121)
// kEmit = false: values only (the adjoint sweep recomputes a frame without counting its residual rows again)
template <bool kEmit> RL_HD void phaseObserve(const Ctx &c, int lane, int frame, double &loss) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  const int nq = H[H_NQ];
  const long row0 = (long)frame * H[H_ROWS_FRAME];
  for (int ci = lane; ci < nq; ci += 32) {
    const double *cd = c.D + H[H_D_CTRL] + ci * CD_STRIDE;
    const double q = s[H[R_Q] + ci], w = dc[DC_W_JLIMIT];
    if (kEmit) {
      emit(c, row0 + 2 * ci, relu(q - cd[CD_UPPER]) * w, loss);
      emit(c, row0 + 2 * ci + 1, relu(cd[CD_LOWER] - q) * w, loss);
    }
    if (H[H_NLAYERS]) c.x[ci] = q;
  }
  if (lane == 31 && H[H_NLAYERS]) {
    const PoseV obj = ldp(s + H[R_POSE] + H[H_OBJ] * 7);
    double *x = c.x + nq;
    x[0] = obj.t.x * 10.0, x[1] = obj.t.y * 10.0, x[2] = obj.t.z * 10.0;
    st3(x + 3, qrot(obj.r, V3{1, 0, 0}));
    st3(x + 6, qrot(obj.r, V3{0, 1, 0}));
    st3(x + 9, ld3(s + H[R_POSE] + H[H_PALM] * 7) - obj.t);
  }
  if (lane == 0 && frame == 0) {
    // first_link_pose (dexsyn_assemble.h rollout(): MoveGoal / RelativeOrientationGoal compare against it)
    for (int e = 0; e < 7; e++) s[H[R_FIRST] + e] = s[H[R_POSE] + H[H_OBJ] * 7 + e];
  }
}

// hidden layer: activity regularisation goal, then tanh (dexenv_grasp.h:137-146)
RL_HD void phaseHidden(const Ctx &c, int lane, int frame, double &loss) {
  const int32_t *I = c.I, *H = c.H;
  const double reg = dconst(c)[DC_W_REG];
  const long row0 = (long)frame * H[H_ROWS_FRAME] + H[H_ROW_ACT];
  for (int j = lane; j < H[H_NHID]; j += 32) {
    const double h = c.h[j];
    emit(c, row0 + j, h * reg, loss);
    c.a[j] = tanh(h);
  }
}

// command regularisation and tanh of the policy output (dexenv_grasp.h:186-199)
RL_HD void phaseControl(const Ctx &c, int lane, int frame, double &loss) {
  const int32_t *I = c.I, *H = c.H;
  const double creg = dconst(c)[DC_W_CREG];
  const long row0 = (long)frame * H[H_ROWS_FRAME] + H[H_ROW_CMD];
  const int nq = H[H_NQ];
  for (int i = lane; i < H[H_NOUT]; i += 32) {
    const double o = c.out[i];
    emit(c, row0 + i, o * creg, loss);
    if (i < nq) c.s[H[R_VEL] + i] = tanh(o);
  }
}

// shape penalty of a contact point (dexlearn.h:159-217): residual rows and, for the adjoint, the gradient with
// respect to the point in link coordinates
template <bool kReverse>
RL_HD V3 shapeRows(const Ctx &c, const int32_t *si, const double *sd, const V3 &p, long row, double &loss) {
  const double f = dconst(c)[DC_W_SHAPE];
  V3 g{0, 0, 0};
  if (si[EI_KIND] == 1) {
    const double half = sd[ED_LENGTH] * 0.5, radius = sd[ED_RADIUS], length = sd[ED_LENGTH];
    {
      // f * (relu(pz - l/2) + relu(-pz - l/2))
      const double a1 = p.z - half, a2 = -p.z - half;
      const double v = f * (relu(a1) + relu(a2));
      if (!kReverse) emit(c, row, v, loss);
      else {
        const double gs = seedOf(c, row, v) * f;  // reverse_mul: d(relu sum) = dx * f
        if (a1 >= 0.0) g.z += gs;                 // reverse_relu: a >= 0
        if (a2 >= 0.0) g.z -= gs;
      }
    }
    for (int i = 0; i < 8; i++) {
      const double angle = i * 3.14159265358979323846 / 8;
      const double sn = sin(angle), cs = cos(angle);
      const double v0 = p.x * sn + p.y * cs;
      const double a1 = v0 - radius, a2 = -v0 - length;
      const double v = f * (relu(a1) + relu(a2));
      if (!kReverse) emit(c, row + 1 + i, v, loss);
      else {
        const double gs = seedOf(c, row + 1 + i, v) * f;
        double gv = 0;
        if (a1 >= 0.0) gv += gs;
        if (a2 >= 0.0) gv -= gs;
        g.x += gv * sn;
        g.y += gv * cs;
      }
    }
    return g;
  }
  const double *pl = c.D + c.H[H_D_PLANES] + si[EI_PLANE0] * 4;
  for (int i = 0; i < si[EI_NPLANES]; i++, pl += 4) {
    const V3 n = ld3(pl);
    const double dist = dot(n, p) + pl[3];
    const double v = f * relu(dist);
    if (!kReverse) emit(c, row + i, v, loss);
    else if (dist >= 0.0) g = g + n * (seedOf(c, row + i, v) * f);  // reverse_mul, reverse_relu, reverse_dot_fg_vec3
  }
  return g;
}

// which link / shape / scratch slot does lane item (end effector e, side) work on?
struct Side {
  int e, side, link, track;
  const int32_t *si;
  const double *sd;
  long row_shape;  // first shape-penalty row of this side
  long row_ee;     // first row of the end effector
};
RL_HD Side sideOf(const Ctx &c, int item, int frame) {
  const int32_t *I = c.I, *H = c.H;
  Side sd;
  sd.e = item >> 1;
  sd.side = item & 1;
  const int32_t *ei = I + H[H_I_EE] + sd.e * EI_STRIDE;
  const int32_t *oi = I + H[H_I_OBJSHAPE];
  sd.si = sd.side ? ei : oi;
  sd.sd = sd.side ? c.D + H[H_D_EE] + sd.e * ED_STRIDE : c.D + H[H_D_OBJSHAPE];
  sd.link = sd.si[EI_LINK];
  sd.track = sd.si[EI_TRACK];
  sd.row_ee = (long)frame * H[H_ROWS_FRAME] + ei[EI_ROW];
  sd.row_shape = sd.row_ee + 9 + (sd.side ? oi[EI_NROWS] : 0);
  return sd;
}

// contacts, step 1 (dexlearn.h:250-283): contact point of one side, its regularisation and shape penalty
template <bool kEmit> RL_HD void phaseContact1(const Ctx &c, int lane, int frame, double &loss) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  const int nq = H[H_NQ];
  for (int item = lane; item < 2 * H[H_NEE]; item += 32) {
    const Side sd = sideOf(c, item, frame);
    const double *cp = c.out + nq + sd.e * H[H_CDIMS] + sd.side * 3;
    const V3 cs = ld3(cp) * 0.1;
    const V3 reg = cs * dc[DC_W_CPR];
    if (kEmit) {
      emit(c, sd.row_ee + sd.side * 3 + 0, reg.x, loss);
      emit(c, sd.row_ee + sd.side * 3 + 1, reg.y, loss);
      emit(c, sd.row_ee + sd.side * 3 + 2, reg.z, loss);
    }
    const PoseV pose = ldp(s + H[R_POSE] + sd.link * 7);
    const V3 pt = pmulv(pose, ld3(sd.sd + ED_CENTER)) + cs;
    st3(s + H[R_EE] + sd.e * EE_STRIDE + (sd.side ? EE_C2 : EE_C1), pt);
    if (kEmit) {
      const V3 local = pmulv(pinv(pose), pt);
      shapeRows<false>(c, sd.si, sd.sd, local, sd.row_shape, loss);
    }
  }
  // velocity-controller commands for the next frame (dexenv_grasp.h:200-219): coupling, scaling
  for (int ci = lane; ci < nq; ci += 32) {
    const int src = I[H[H_I_CTRL] + ci * CI_STRIDE + CI_SRC];
    s[H[R_CMD] + ci] = s[H[R_VEL] + src] * c.D[H[H_D_CTRL] + ci * CD_STRIDE + CD_SCALE];
  }
}

// PhysicsSimulator::_velocity (physics5.h:106-112)
RL_HD V3 pointVelocity(const PoseV &pose, const PoseV &prev, const V3 &p, double inv_dt) {
  const V3 moved = pmulv(prev, pmulv(pinv(pose), p));
  return (p - moved) * inv_dt;
}

// contacts, step 2 (dexlearn.h:375-383): velocity of the contact midpoint as seen from each link
RL_HD void phaseContact2(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  for (int item = lane; item < 2 * H[H_NEE]; item += 32) {
    const Side sd = sideOf(c, item, frame);
    double *ee = s + H[R_EE] + sd.e * EE_STRIDE;
    const V3 p = (ld3(ee + EE_C1) + ld3(ee + EE_C2)) * 0.5;
    const V3 v = pointVelocity(ldp(s + H[R_POSE] + sd.link * 7), ldp(s + H[R_PREV] + sd.track * 7), p,
                               dconst(c)[DC_INV_DT]);
    st3(ee + (sd.side ? EE_V2 : EE_V1), v);
  }
}

// contacts, step 3 (dexlearn.h:285-312,364-390): force, distance (x) force, slip (x) force, impulse
RL_HD void phaseContact3(const Ctx &c, int lane, int frame, double &loss) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  for (int e = lane; e < H[H_NEE]; e += 32) {
    double *ee = s + H[R_EE] + e * EE_STRIDE;
    const int32_t *ei = I + H[H_I_EE] + e * EI_STRIDE;
    const int32_t *oi = I + H[H_I_OBJSHAPE];
    const long row_ee = (long)frame * H[H_ROWS_FRAME] + ei[EI_ROW];
    const long row_t = row_ee + 9 + oi[EI_NROWS] + ei[EI_NROWS];
    const double *cp = c.out + H[H_NQ] + e * H[H_CDIMS];
    const V3 force = ld3(cp + 6) * 0.1;
    const V3 freg = force * dc[DC_W_CFR];
    emit(c, row_ee + 6, freg.x, loss);
    emit(c, row_ee + 7, freg.y, loss);
    emit(c, row_ee + 8, freg.z, loss);
    const V3 c1 = ld3(ee + EE_C1), c2 = ld3(ee + EE_C2);
    const V3 d = (c2 - c1) * dc[DC_W_CDP];
    const V3 dv = (ld3(ee + EE_V2) - ld3(ee + EE_V1)) * dc[DC_W_CSP];
    const double f[3] = {force.x, force.y, force.z};
    for (int k = 0; k < 3; k++) {
      emit(c, row_t + 3 * k + 0, d.x * f[k], loss);
      emit(c, row_t + 3 * k + 1, d.y * f[k], loss);
      emit(c, row_t + 3 * k + 2, d.z * f[k], loss);
      emit(c, row_t + 9 + 3 * k + 0, dv.x * f[k], loss);
      emit(c, row_t + 9 + 3 * k + 1, dv.y * f[k], loss);
      emit(c, row_t + 9 + 3 * k + 2, dv.z * f[k], loss);
    }
    // applyForce (physics5.h:114-121,672-675): the momenta take these at the start of the next frame
    const V3 impulse = force * dc[DC_DT];
    st3(ee + EE_IMP, impulse);
    st3(ee + EE_IMP + 3, cross(c1 - ld3(s + H[R_CG]), impulse));
  }
}

// terminal goals: MoveGoal (goals.h:173-191) and RelativeOrientationGoal (goals.h:71-89)
struct Terminal {
  V3 r_move, r_ori;
};
RL_HD Terminal terminalResiduals(const Ctx &c) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  const PoseV last = ldp(c.s + H[R_POSE] + H[H_OBJ] * 7), first = ldp(c.s + H[R_FIRST]);
  Terminal t;
  t.r_move = ((last.t - first.t) - ld3(dc + DC_GOAL)) * dc[DC_W_MOVE];
  const V3 res = qresidual(qmul(qinv(first.r), last.r));
  t.r_ori = (V3{0, 0, 0} - res) * dc[DC_W_ORI];
  return t;
}
RL_HD void phaseTerminal(const Ctx &c, int lane, double &loss) {
  if (lane != 0) return;
  const Terminal t = terminalResiduals(c);
  const long row = (long)c.frames * c.H[H_ROWS_FRAME];
  emit(c, row + 0, t.r_move.x, loss);
  emit(c, row + 1, t.r_move.y, loss);
  emit(c, row + 2, t.r_move.z, loss);
  emit(c, row + 3, t.r_ori.x, loss);
  emit(c, row + 4, t.r_ori.y, loss);
  emit(c, row + 5, t.r_ori.z, loss);
}

// initial state of a rollout (dexsyn_assemble.h rollout(): start configuration, simulator.init, the
// environment's random object offset and yaw as per-rollout parameters {px, py, yaw})
RL_HD void phaseInit(const Ctx &c, int lane, const double *params) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  const int nq = H[H_NQ];
  for (int e = lane; e < H[H_NSTATIC] * 7; e += 32) s[H[R_POSE] + e] = c.D[H[H_D_STATIC] + e];
  // local rotation of a fixed joint: its origin's, for good
  for (int k = lane; k < H[H_NJ]; k += 32)
    if (I[H[H_I_JOINT] + k * JI_STRIDE + JI_TYPE] != dexsyn::JOINT_REVOLUTE)
      st4(s + H[R_LQ] + k * 4, ld4(c.D + H[H_D_JOINT] + k * JD_STRIDE + JD_ORIGIN + 3));
  for (int ci = lane; ci < nq; ci += 32) {
    const double q = c.D[H[H_D_CTRL] + ci * CD_STRIDE + CD_Q0];
    s[H[R_Q] + ci] = q;
    s[H[R_CMD] + ci] = 0.0;
    const int k = I[H[H_I_CTRL] + ci * CI_STRIDE + CI_JOINT];
    const double *jd = c.D + H[H_D_JOINT] + k * JD_STRIDE;
    st4(s + H[R_LQ] + k * 4, qmul(ld4(jd + JD_ORIGIN + 3), qangleaxis(q, ld3(jd + JD_AXIS))));
  }
  if (lane == 31) {
    const PoseV init = ldp(dc + DC_OBJ_INIT);
    stp(s + H[R_POSE] + H[H_OBJ] * 7, init);
    double *body = s + H[R_BODY];
    st3(body + B_POS, init.t + V3{params[0], params[1], 0.0});
    st4(body + B_ORI, qmul(qangleaxis(params[2], V3{0, 0, 1}), init.r));
    st3(body + B_LM, V3{0, 0, 0});
    st3(body + B_AM, V3{0, 0, 0});
  }
}

// ================================================================================== adjoint phases

// start of an adjoint frame: link-pose adjoints start from what the next frame's slip velocities left for
// the tracked links (their "previous" pose is this frame's pose)
RL_HD void bphaseBegin(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  // every other link-pose adjoint is zero: its consumer (bphaseChainWrench, the body lane of bphaseSim) cleared it
  for (int e = lane; e < H[H_NTRACK] * 6; e += 32) {
    const int t = e / 6, k = I[H[H_I_TRACK] + t];
    s[H[R_GPOSE] + k * 6 + (e - t * 6)] = s[H[R_GPREV] + e];
    s[H[R_GPREV] + e] = 0.0;  // refilled by this frame's slip velocities
  }
}
RL_HD void bphaseBegin2(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  // first and last frame only: the terminal goals look at the object pose of both
  if (lane == 31) {
    if (frame == c.frames - 1) {
      // reverse of the terminal goals: reverse_mul_fg_vec3_s, reverse_sub_fg_vec3 (second operand negated),
      // reverse_fg_quat_residual (identity), reverse_mul_fg_quat (da = dx, db = qrot(qinv(va), dx)),
      // reverse_fg_quat_inverse (negation)
      const double *dc = dconst(c);
      const Terminal t = terminalResiduals(c);
      const long row = (long)c.frames * H[H_ROWS_FRAME];
      const V3 sm{seedOf(c, row + 0, t.r_move.x), seedOf(c, row + 1, t.r_move.y), seedOf(c, row + 2, t.r_move.z)};
      const V3 so{seedOf(c, row + 3, t.r_ori.x), seedOf(c, row + 4, t.r_ori.y), seedOf(c, row + 5, t.r_ori.z)};
      const V3 g_moved = sm * dc[DC_W_MOVE];
      const V3 g_qm = -(so * dc[DC_W_ORI]);
      const PoseV first = ldp(s + H[R_FIRST]);
      TwistV g_last{g_moved, qrot(first.r, g_qm)};  // qinv(va) with va = qinv(first.r)
      TwistV g_first{-g_moved, -g_qm};
      addt(s + H[R_GPOSE] + H[H_OBJ] * 6, g_last);
      stt(s + H[R_GFIRST], g_first);
    }
    if (frame == 0) addt(s + H[R_GPOSE] + H[H_OBJ] * 6, ldt(s + H[R_GFIRST]));
  }
}

// adjoint of contact step 3 and of the impulse
RL_HD void bphaseContact3(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  const V3 g_lm = ld3(s + H[R_GBODY] + GB_LM), g_am = ld3(s + H[R_GBODY] + GB_AM);
  for (int e = lane; e < H[H_NEE]; e += 32) {
    double *ee = s + H[R_EE] + e * EE_STRIDE;
    const int32_t *ei = I + H[H_I_EE] + e * EI_STRIDE;
    const int32_t *oi = I + H[H_I_OBJSHAPE];
    const long row_ee = (long)frame * H[H_ROWS_FRAME] + ei[EI_ROW];
    const long row_t = row_ee + 9 + oi[EI_NROWS] + ei[EI_NROWS];
    const int o = H[H_NQ] + e * H[H_CDIMS];
    const V3 force = ld3(c.out + o + 6) * 0.1;
    const V3 c1 = ld3(ee + EE_C1), c2 = ld3(ee + EE_C2);
    const V3 d = (c2 - c1) * dc[DC_W_CDP];
    const V3 dv = (ld3(ee + EE_V2) - ld3(ee + EE_V1)) * dc[DC_W_CSP];
    const double f[3] = {force.x, force.y, force.z};
    double gf[3];
    V3 g_d{0, 0, 0}, g_dv{0, 0, 0};
    for (int k = 0; k < 3; k++) {
      // reverse_mul_fg_vec3_s: da = dx * b, db = dot(dx, a)
      const V3 sd{seedOf(c, row_t + 3 * k, d.x * f[k]), seedOf(c, row_t + 3 * k + 1, d.y * f[k]),
                  seedOf(c, row_t + 3 * k + 2, d.z * f[k])};
      const V3 sv{seedOf(c, row_t + 9 + 3 * k, dv.x * f[k]), seedOf(c, row_t + 9 + 3 * k + 1, dv.y * f[k]),
                  seedOf(c, row_t + 9 + 3 * k + 2, dv.z * f[k])};
      g_d = g_d + sd * f[k];
      g_dv = g_dv + sv * f[k];
      gf[k] = dot(sd, d) + dot(sv, dv);
    }
    // impulse = force * dt; lm += impulse; am += cross(c1 - cg, impulse)   (reverse_cross_fg_vec3)
    const V3 impulse = force * dc[DC_DT];
    const V3 lever = c1 - ld3(s + H[R_CG]);
    const V3 g_lever = cross(impulse, g_am);
    const V3 g_impulse = g_lm + cross(g_am, lever);
    const V3 freg = force * dc[DC_W_CFR];
    const V3 sf{seedOf(c, row_ee + 6, freg.x), seedOf(c, row_ee + 7, freg.y), seedOf(c, row_ee + 8, freg.z)};
    const V3 g_force = g_impulse * dc[DC_DT] + sf * dc[DC_W_CFR];
    c.gout[o + 6] = (g_force.x + gf[0]) * 0.1;
    c.gout[o + 7] = (g_force.y + gf[1]) * 0.1;
    c.gout[o + 8] = (g_force.z + gf[2]) * 0.1;
    const V3 g_c = g_d * dc[DC_W_CDP];
    st3(ee + EE_GC1, g_lever - g_c);
    st3(ee + EE_GC2, g_c);
    const V3 g_v = g_dv * dc[DC_W_CSP];
    st3(ee + EE_GV2, g_v);
    st3(ee + EE_GV1, -g_v);
    st3(ee + EE_GP, -g_lever);  // adjoint of the body centre from this end effector (summed by the body lane)
  }
}

// adjoint of the point velocity: to the point, the link pose and the previous link pose
RL_HD void bphaseContact2(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  const double inv_dt = dconst(c)[DC_INV_DT];
  for (int item = lane; item < 2 * H[H_NEE]; item += 32) {
    const Side sd = sideOf(c, item, frame);
    double *ee = s + H[R_EE] + sd.e * EE_STRIDE;
    const V3 p = (ld3(ee + EE_C1) + ld3(ee + EE_C2)) * 0.5;
    const PoseV pose = ldp(s + H[R_POSE] + sd.link * 7), prev = ldp(s + H[R_PREV] + sd.track * 7);
    const PoseV inv = pinv(pose);
    const V3 pl = pmulv(inv, p);
    const V3 g_v = ld3(ee + (sd.side ? EE_GV2 : EE_GV1));
    const V3 g_diff = g_v * inv_dt;                    // reverse_mul_fg_vec3_s (1/dt is a constant)
    const V3 g_moved = -g_diff;                        // reverse_sub_fg_vec3
    // moved = prev * pl: reverse_mul_fg_pose_vec3: da = (dx, cross(qrot(a.r, b), dx)), db = qrot(qinv(a.r), dx)
    const TwistV g_prev{g_moved, cross(qrot(prev.r, pl), g_moved)};
    const V3 g_pl = qrot(qinv(prev.r), g_moved);
    // pl = inv * p
    const TwistV g_inv{g_pl, cross(qrot(inv.r, p), g_pl)};
    const V3 g_p = g_diff + qrot(pose.r, g_pl);        // qinv(inv.r) = pose.r
    const TwistV g_pose = pinvReverse(pose, inv, g_inv);
    // side 1 owns its link; the object link is shared by all end effectors -> per end effector, summed later
    if (sd.side) {
      addt(s + H[R_GPOSE] + sd.link * 6, g_pose);
      addt(s + H[R_GPREV] + sd.track * 6, g_prev);
      st3(ee + EE_GP + 3, g_p);
    } else {
      stt(s + H[R_GOBJ] + sd.e * 6, g_pose);
      stt(s + H[R_GPREVOBJ] + sd.e * 6, g_prev);
      // EE_GP holds the body-centre adjoint until the body lane has read it; the midpoint adjoint of side 0
      // travels in EE_GV1 (consumed above)
      st3(ee + EE_GV1, g_p);
    }
  }
}

// adjoint of contact step 1
RL_HD void bphaseContact1(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  const int nq = H[H_NQ];
  double unused = 0;
  for (int item = lane; item < 2 * H[H_NEE]; item += 32) {
    const Side sd = sideOf(c, item, frame);
    double *ee = s + H[R_EE] + sd.e * EE_STRIDE;
    const int o = nq + sd.e * H[H_CDIMS] + sd.side * 3;
    const V3 cs = ld3(c.out + o) * 0.1;
    const PoseV pose = ldp(s + H[R_POSE] + sd.link * 7);
    const V3 center = ld3(sd.sd + ED_CENTER);
    const V3 pt = ld3(ee + (sd.side ? EE_C2 : EE_C1));
    const PoseV inv = pinv(pose);
    const V3 local = pmulv(inv, pt);
    // midpoint p = (c1 + c2) * 0.5 feeds both velocities
    V3 g_pt = ld3(ee + (sd.side ? EE_GC2 : EE_GC1)) + (ld3(ee + EE_GV1) + ld3(ee + EE_GP + 3)) * 0.5;
    // shape penalty: local = inv * pt
    const V3 g_local = shapeRows<true>(c, sd.si, sd.sd, local, sd.row_shape, unused);
    const TwistV g_inv{g_local, cross(qrot(inv.r, pt), g_local)};
    g_pt = g_pt + qrot(pose.r, g_local);
    TwistV g_pose = pinvReverse(pose, inv, g_inv);
    // pt = pose * center + cs   (centre is a constant)
    g_pose.t = g_pose.t + g_pt;
    g_pose.r = g_pose.r + cross(qrot(pose.r, center), g_pt);
    const V3 reg = cs * dc[DC_W_CPR];
    const V3 sr{seedOf(c, sd.row_ee + sd.side * 3, reg.x), seedOf(c, sd.row_ee + sd.side * 3 + 1, reg.y),
                seedOf(c, sd.row_ee + sd.side * 3 + 2, reg.z)};
    const V3 g_cs = g_pt + sr * dc[DC_W_CPR];
    c.gout[o + 0] = g_cs.x * 0.1;
    c.gout[o + 1] = g_cs.y * 0.1;
    c.gout[o + 2] = g_cs.z * 0.1;
    if (sd.side) addt(s + H[R_GPOSE] + sd.link * 6, g_pose);
    else addt(s + H[R_GOBJ] + sd.e * 6, g_pose);
  }
}

// adjoint of the control step: commands -> tanh -> policy output, plus the command regularisation rows.
// Runs after the contact phases have written the contact-parameter part of gout.
RL_HD void bphaseControl(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  const double creg = dconst(c)[DC_W_CREG];
  double *s = c.s;
  const int nq = H[H_NQ];
  const long row0 = (long)frame * H[H_ROWS_FRAME] + H[H_ROW_CMD];
  for (int i = lane; i < H[H_NOUT]; i += 32) {
    const double o = c.out[i];
    const double gs = seedOf(c, row0 + i, o * creg) * creg;
    if (i < nq) {
      // the command of joint i and, with the J1 := J2 coupling, of the joint it also drives (dexenv_grasp.h:200-219)
      double g_vel = 0;
      const int32_t *cidx = I + H[H_I_CTRL] + i * CI_STRIDE;
      if (cidx[CI_SRC] == i) g_vel += s[H[R_GCMD] + i] * c.D[H[H_D_CTRL] + i * CD_STRIDE + CD_SCALE];
      if (cidx[CI_DST] >= 0) g_vel += s[H[R_GCMD] + cidx[CI_DST]] * c.D[H[H_D_CTRL] + cidx[CI_DST] * CD_STRIDE + CD_SCALE];
      const double v = s[H[R_VEL] + i];
      c.gout[i] = g_vel * (1.0 - v * v) + gs;  // prepare_tanh saves 1 - x^2
    } else {
      c.gout[i] += gs;
    }
  }
}

// adjoint of tanh and of the activity regularisation rows of the hidden layer; ga arrives in c.gh
RL_HD void bphaseHidden(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  const double reg = dconst(c)[DC_W_REG];
  const long row0 = (long)frame * H[H_ROWS_FRAME] + H[H_ROW_ACT];
  for (int j = lane; j < H[H_NHID]; j += 32) {
    const double a = c.a[j];
    c.gh[j] = c.gh[j] * (1.0 - a * a) + seedOf(c, row0 + j, c.h[j] * reg) * reg;
  }
}

// adjoint of the observation and of the joint-limit rows; gathers what the end effectors left for the object
RL_HD void bphaseObserve(const Ctx &c, int lane, int frame) {
  const int32_t *I = c.I, *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  const int nq = H[H_NQ];
  const long row0 = (long)frame * H[H_ROWS_FRAME];
  const bool policy = H[H_NLAYERS] != 0;
  for (int ci = lane; ci < nq; ci += 32) {
    const double *cd = c.D + H[H_D_CTRL] + ci * CD_STRIDE;
    const double q = s[H[R_Q] + ci], w = dc[DC_W_JLIMIT];
    const double au = q - cd[CD_UPPER], al = cd[CD_LOWER] - q;
    double g = policy ? c.gx[ci] : 0.0;
    if (au >= 0.0) g += seedOf(c, row0 + 2 * ci, relu(au) * w) * w;
    if (al >= 0.0) g -= seedOf(c, row0 + 2 * ci + 1, relu(al) * w) * w;
    s[H[R_GQ] + ci] += g;
  }
  if (lane == 31 && policy) {
    const PoseV obj = ldp(s + H[R_POSE] + H[H_OBJ] * 7);
    const double *gx = c.gx + nq;
    TwistV g_obj{ld3(gx) * 10.0, V3{0, 0, 0}};
    // reverse_mul_fg_quat_vec3 with a constant vector: da = cross(qrot(q, e), dx)
    g_obj.r = cross(qrot(obj.r, V3{1, 0, 0}), ld3(gx + 3)) + cross(qrot(obj.r, V3{0, 1, 0}), ld3(gx + 6));
    const V3 g_rel = ld3(gx + 9);
    g_obj.t = g_obj.t - g_rel;
    add3(s + H[R_GPOSE] + H[H_PALM] * 6, g_rel);
    addt(s + H[R_GPOSE] + H[H_OBJ] * 6, g_obj);
  }
}
// what the end effectors left for the object link and for its previous pose (one slot per end effector, because
// all of them touch the same link): summed component by component, lanes 0..5 the pose adjoint, 6..11 the previous
RL_HD void bphaseObjectGather(const Ctx &c, int lane) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  if (lane >= 12 || !H[H_NLAYERS]) return;
  const int comp = lane < 6 ? lane : lane - 6;
  const double *src = s + (lane < 6 ? H[R_GOBJ] : H[R_GPREVOBJ]) + comp;
  double sum = 0.0;
  for (int e = 0; e < H[H_NEE]; e++) sum += src[e * 6];
  if (lane < 6) s[H[R_GPOSE] + H[H_OBJ] * 6 + comp] += sum;
  else s[H[R_GPREV] + I[H[H_I_OBJSHAPE] + EI_TRACK] * 6 + comp] += sum;
}

// adjoint of the forward kinematics.  What a joint hands to its parent is (g.t, cross(p_k - p_parent, g.t) + g.r)
// (reverse_mul_fg_pose: da = (dx.t, cross(qrot(a.r, b.t), dx.t) + dx.r)): force and torque of a wrench moved from the
// joint's origin to the parent's.  Referred to the world origin instead, moving is the identity, so the total a
// joint sees is a plain sum over its subtree; the joint angle takes the torque about its own origin,
// M_k - cross(p_k, F_k), through reverse_fg_pose_angle_axis_pose: dangle = dot(qrot(qinv(parent.r), torque), axis).
// Same terms as a joint-by-joint walk, summed in a different order.
// the adjoint of the forward kinematics as warp scans, one round (deepest first): wrench of every joint about the
// world origin plus what the segments hanging off it collected, suffix sums inside the segments, joint angles
template <int N, class W> RL_HD void fkRoundReverse(const Ctx &c, int round, int lane0) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  const int steps = I[H[H_I_ROUNDSTEPS] + round];
  RoundLane rl[N];
  TwistV w[N], o[N];
  V3 p[N];
  for (int i = 0; i < N; i++) {
    rl[i] = roundLane(c, round, lane0 + i);
    w[i] = TwistV{{0, 0, 0}, {0, 0, 0}};
    p[i] = V3{0, 0, 0};
    if (rl[i].k < 0) continue;
    const int k = rl[i].k;
    double *gp = s + H[R_GPOSE] + k * 6;
    const TwistV g = ldt(gp);
    stt(gp, TwistV{{0, 0, 0}, {0, 0, 0}});
    p[i] = ld3(s + H[R_POSE] + k * 7);
    w[i] = TwistV{g.t, g.r + cross(p[i], g.t)};
    const int32_t *ji = I + H[H_I_JOINT] + k * JI_STRIDE;
    for (int ch = 0; ch < ji[JI_NCHILD]; ch++) {
      const TwistV gc = ldt(s + H[R_GC] + I[ji[JI_CHILD0] + ch] * 6);
      w[i].t = w[i].t + gc.t;
      w[i].r = w[i].r + gc.r;
    }
  }
  for (int st = 0, d = 1; st < steps; st++, d <<= 1) {
    W::down(w, d, o);
    for (int i = 0; i < N; i++)
      if (rl[i].j + d < rl[i].len) {
        w[i].t = w[i].t + o[i].t;
        w[i].r = w[i].r + o[i].r;
      }
  }
  for (int i = 0; i < N; i++) {
    if (rl[i].k < 0) continue;
    const int k = rl[i].k;
    if (rl[i].j == 0) stt(s + H[R_GC] + k * 6, w[i]);  // the segment's total, for the joint it hangs off
    const int32_t *ji = I + H[H_I_JOINT] + k * JI_STRIDE;
    if (ji[JI_TYPE] != dexsyn::JOINT_REVOLUTE) continue;
    const V3 torque = w[i].r - cross(p[i], w[i].t);
    const Q4 rpar = ld4(s + H[R_POSE] + ji[JI_PARENT] * 7 + 3);
    s[H[R_GQ] + ji[JI_CTRL]] += dot(qrot(qinv(rpar), torque), ld3(c.D + H[H_D_JOINT] + k * JD_STRIDE + JD_CAXIS));
  }
}
// pose adjoints of the links outside the segments (statics: the floor collects contact adjoints nobody reads);
// the object's is consumed and cleared by the body lane of bphaseSim
RL_HD void fkClearStatics(const Ctx &c, int lane) {
  const int32_t *H = c.H;
  for (int k = lane; k < H[H_KSEG0]; k += 32)
    if (k != H[H_OBJ]) stt(c.s + H[R_GPOSE] + k * 6, TwistV{{0, 0, 0}, {0, 0, 0}});
}

// adjoint of the simulator step, joints: q = q_prev + cmd_prev * dt   (reverse_add, reverse_mul with a constant)
RL_HD void bphaseSim(const Ctx &c, int lane) {
  const int32_t *H = c.H;
  const double dt = dconst(c)[DC_DT];
  for (int ci = lane; ci < H[H_NQ]; ci += 32) c.s[H[R_GCMD] + ci] = c.s[H[R_GQ] + ci] * dt;
}
// adjoint of the simulator step, body (one lane per rollout)
RL_HD void bphaseBody(const Ctx &c) {
  const int32_t *H = c.H;
  const double *dc = dconst(c);
  double *s = c.s;
  BodyStep b;
  bodyForward(dc, s + H[R_BODY0], b);
  const double dt = dc[DC_DT];
  double *gb = s + H[R_GBODY];
  const V3 g_lm_out = ld3(gb + GB_LM), g_am3 = ld3(gb + GB_AM);
  V3 g_pos2 = ld3(gb + GB_POS), g_ori2 = ld3(gb + GB_ORI);
  V3 g_cg{0, 0, 0};
  for (int e = 0; e < H[H_NEE]; e++) g_cg = g_cg + ld3(s + H[R_EE] + e * EE_STRIDE + EE_GP);
  // object link pose D = C * R(fo), C = origin * T(fp); fp, fo = translation / orientation of Bp
  const TwistV g_D = ldt(s + H[R_GPOSE] + H[H_OBJ] * 6);
  stt(s + H[R_GPOSE] + H[H_OBJ] * 6, TwistV{{0, 0, 0}, {0, 0, 0}});
  const V3 g_C_t = g_D.t;                                    // reverse_mul_fg_pose da.t (b.t = 0: no lever arm)
  const V3 g_fo = qrot(qinv(b.C.r), g_D.r);                  // db.r -> reverse_fg_orientation_pose
  const PoseV org = ldp(dc + DC_OBJ_ORIGIN);
  const V3 g_fp = qrot(qinv(org.r), g_C_t);                  // db.t -> reverse_fg_translation_pose
  // Bp = A * R(ori2), A = anchor * T(pos2)
  const V3 g_A_t = g_fp;                                     // reverse_fg_pose_translation, da.t
  g_ori2 = g_ori2 + qrot(qinv(b.A.r), g_fo);                 // reverse_fg_pose_orientation, db.r
  const PoseV ia = ldp(dc + DC_ANCHOR);
  g_pos2 = g_pos2 + qrot(qinv(ia.r), g_A_t);
  // am3 = ori2 * lam
  g_ori2 = g_ori2 + cross(b.am3, g_am3);
  V3 g_lam = qrot(qinv(b.ori2), g_am3);
  // ori2 = ori (+) rot   (reverse_add_fg_quat_vec3: both get dx)
  V3 g_ori = g_ori2;
  V3 g_rot = g_ori2;
  // pos2 = pos1 - cross(rot, rc)
  const V3 g_pos1 = g_pos2;
  const V3 g_cr = -g_pos2;
  g_rot = g_rot + cross(b.rc, g_cr);                         // reverse_cross_fg_vec3: da = cross(vb, dx)
  const V3 g_rc = cross(g_cr, b.rot);                        //                        db = cross(dx, va)
  g_ori = g_ori + cross(b.rc, g_rc);                         // rc = ori * center
  const V3 g_av = g_rot * dt;
  g_ori = g_ori + cross(b.av, g_av);                         // av = ori * w
  const V3 g_w = qrot(qinv(b.ori), g_av);
  {
    const double *m = dc + DC_IINV;                          // reverse_mul_fg_mat3_vec3: db[col] += dx[row] * a[row][col]
    double dx[3] = {g_w.x, g_w.y, g_w.z}, db[3] = {0, 0, 0};
    for (int row = 0; row < 3; row++)
      for (int col = 0; col < 3; col++) db[col] += dx[row] * m[row * 3 + col];
    g_lam = g_lam + V3{db[0], db[1], db[2]};
  }
  V3 g_pos = g_pos1;
  const V3 g_lv = g_pos1 * dt;
  const V3 g_lm2 = g_lm_out + g_lv * dc[DC_INV_MASS];
  // lam = qinv(ori) * am2
  const V3 g_qi = cross(b.lam, g_lam);
  const V3 g_am2 = qrot(b.ori, g_lam);
  g_ori = g_ori - g_qi;                                      // reverse_fg_quat_inverse
  // cg = pos + ori * center
  g_pos = g_pos + g_cg;
  g_ori = g_ori + cross(b.cg - b.pos, g_cg);
  st3(gb + GB_POS, g_pos);
  st3(gb + GB_ORI, g_ori);
  st3(gb + GB_LM, g_lm2 * dc[DC_DAMP_LIN]);
  st3(gb + GB_AM, g_am2 * dc[DC_DAMP_ANG]);
}

// zero the carried adjoints before the adjoint sweep
RL_HD void bphaseInit(const Ctx &c, int lane) {
  const int32_t *I = c.I, *H = c.H;
  double *s = c.s;
  for (int e = lane; e < H[H_NQ]; e += 32) s[H[R_GQ] + e] = 0.0, s[H[R_GCMD] + e] = 0.0;
  if (lane == 0) s[H[R_GQ] + H[H_NQ]] = 0.0;
  for (int e = lane; e < GB_COUNT; e += 32) s[H[R_GBODY] + e] = 0.0;
  for (int e = lane; e < H[H_NTRACK] * 6; e += 32) s[H[R_GPREV] + e] = 0.0;
  for (int e = lane; e < 6; e += 32) s[H[R_GFIRST] + e] = 0.0;
  for (int e = lane; e < H[H_NJ] * 6; e += 32) s[H[R_GPOSE] + e] = 0.0;
}

}  // namespace rollout
