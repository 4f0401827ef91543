// rollout_plan.h -- the re-rolled form of dexopt's rollout objective: one *frame program* plus the
// state that is carried from frame to frame, instead of the reference's flat tape.
//
// The reference unrolls DexLearn::_runBatch (dexopt/src/dexlearn.h:393-422) T times onto one tape whose
// memory image is write-once (tractor/src/core/allocator.cpp:34-42), so its size grows with the horizon.
// Every frame records the same operator sequence; only the carried values differ.  The product's objective
// assembly therefore emits, next to the flat tapes of trx_dexsyn_build, this compact description:
//
//   * the kinematic tree as a joint table cut into serial segments (robot/robotmodel.h:128-148 walks it per frame),
//   * the end effectors with their contact shapes (dexlearn.h:159-217,235-391),
//   * the dynamic body and the integrator constants (dexopt/src/physics5.h:78-101,174-193,334-371),
//   * the policy layer shapes (dexopt/src/neural.h:141-190) and the environment weights
//     (dexopt/src/dexenv_grasp.h:19-48),
//   * the carry map: what crosses a frame boundary (joint positions, commands, body state, the link poses
//     the slip velocities compare against) -- this is the per-frame checkpoint the adjoint recomputes from,
//   * the residual row layout, identical to the goal order of the flat tape (csrc/assembly/dexsyn_assemble.h).
//
// csrc/rollout/rollout_frame.h executes it (hand-written forward and adjoint of one frame).
// Host-only code; the tables are flat arrays so that they can be copied to shared memory as they are.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../assembly/dexsyn_model.h"
#include "../vm_math.cuh"

namespace rollout {

// ---- integer header (indices into Plan::i) -------------------------------------------------------
enum {
  H_NJ = 0,       // joints in the pruned pose table (level order; statics first)
  H_NSTATIC,      // leading joints whose pose is a constant
  H_KSEG0,        // first joint of the segments (statics and floating joints come before)
  H_NSUPER,       // super-levels of the segment tree
  H_NSEG,         // segments
  H_NQ,           // controlled (revolute) joints
  H_NEE,          // end effectors
  H_NTRACK,       // links whose previous-frame pose is carried (end-effector links + object)
  H_NLAYERS,      // 0 = no policy, 1 = linear, 2 = one hidden layer
  H_NIN, H_NHID, H_NOUT,
  H_OBJ,          // pose-table index of the object link
  H_PALM,
  H_CDIMS,        // contact parameters per end effector (dexlearn.h:17)
  H_ROWS_FRAME,   // residual rows per frame
  H_ROW_ACT, H_ROW_CMD, H_ROW_EE0,  // row offsets inside a frame
  H_NSCALAR,      // lane-independent residual rows (weight decay), = number of policy parameters
  H_NX,           // policy parameters
  H_W1, H_B1, H_W2, H_B2,  // offsets of the layers inside x (weights row-major [n_in][n_out], then bias)
  H_CKPT,         // doubles per frame checkpoint
  H_REGION,       // doubles of per-rollout working storage
  // offsets of integer tables
  H_I_SUPER,      // [n_super + 1] first segment of each super-level
  H_I_SEG,        // [n_seg * 2] {first joint (pose-table index), joints}
  H_I_JOINT,      // [nj * JI_STRIDE]
  H_I_CTRL,       // [nq * CI_STRIDE]
  H_I_EE,         // [n_ee * EI_STRIDE]
  H_I_TRACK,      // [n_track] pose-table index
  H_I_OBJSHAPE,   // [EI_STRIDE] (link = object)
  H_NROUND,       // warp rounds of the forward kinematics (below)
  H_I_ROUND,      // [n_round][32 lanes][RD_STRIDE]; offset is a multiple of 4
  H_I_ROUNDSTEPS, // [n_round] scan steps: ceil(log2(longest segment of the round))
  // offsets of double tables
  H_D_JOINT,      // [nj * JD_STRIDE]
  H_D_CTRL,       // [nq * CD_STRIDE]
  H_D_EE,         // [n_ee * ED_STRIDE]
  H_D_OBJSHAPE,   // [ED_STRIDE]
  H_D_PLANES,     // planes {nx, ny, nz, d}
  H_D_CONST,      // DC_* below
  H_D_STATIC,     // [n_static * 7] constant poses
  // offsets inside the per-rollout region (doubles)
  R_POSE, R_PREV, R_LQ, R_GPOSE, R_GC, R_Q, R_CMD, R_VEL, R_GQ, R_GCMD, R_BODY, R_GBODY, R_CG, R_GCG, R_FIRST,
  R_GFIRST, R_GPREV, R_EE, R_GOBJ, R_GPREVOBJ, R_BODY0,
  H_COUNT
};

// per joint, integers
enum { JI_PARENT = 0, JI_TYPE, JI_CTRL, JI_CHILD0, JI_NCHILD, JI_TRACK, JI_STRIDE };
// per joint, doubles: origin pose, axis in the joint frame, axis in the parent frame (R_origin * axis)
enum { JD_ORIGIN = 0, JD_AXIS = 7, JD_CAXIS = 10, JD_STRIDE = 13 };
// per controlled joint
// SRC: policy output whose tanh drives it (J1 := J2 coupling); DST: the other joint this output drives (-1: none)
enum { CI_JOINT = 0, CI_SRC, CI_DST, CI_STRIDE };
enum { CD_LOWER = 0, CD_UPPER, CD_Q0, CD_SCALE, CD_STRIDE };
// Forward kinematics as warp scans: the segments of one super-level are packed side by side into the 32 lanes of a
// *round* (a lane per joint); inside a segment the chain products / sums are prefix scans over the lanes, so a
// serial chain of n joints costs log2(n) steps.  Per lane: joint (-1: idle lane), position inside its segment,
// segment length, parent joint of the segment's first joint.
enum { RD_K = 0, RD_J, RD_LEN, RD_PAR, RD_STRIDE };
// per end effector / shape
enum { EI_LINK = 0, EI_KIND, EI_NPLANES, EI_PLANE0, EI_NROWS, EI_TRACK, EI_ROW, EI_STRIDE };
enum { ED_CENTER = 0, ED_RADIUS = 3, ED_LENGTH, ED_STRIDE };
// per-end-effector scratch in the region
enum { EE_C1 = 0, EE_C2 = 3, EE_V1 = 6, EE_V2 = 9, EE_GC1 = 12, EE_GC2 = 15, EE_GV1 = 18, EE_GV2 = 21, EE_GP = 24,
       EE_IMP = 30 /*impulse (3), angular impulse (3) of the frame's contact force*/, EE_STRIDE = 36 };
// constants
enum {
  DC_DT = 0, DC_INV_DT, DC_GRAV = 2 /*3: gravity * (mass * dt)*/, DC_DAMP_LIN = 5, DC_DAMP_ANG, DC_INV_MASS,
  DC_IINV = 8 /*9*/, DC_CENTER = 17 /*3*/, DC_ANCHOR = 20 /*7*/, DC_OBJ_ORIGIN = 27 /*7*/, DC_W_JLIMIT = 34, DC_W_SHAPE,
  DC_W_CPR, DC_W_CFR, DC_W_CDP, DC_W_CSP, DC_W_REG, DC_W_CREG, DC_GOAL = 42 /*3*/, DC_W_MOVE = 45, DC_W_ORI,
  DC_OBJ_INIT = 47 /*7: object link pose of the initial FK*/, DC_COUNT = 54
};
// body state inside R_BODY / a checkpoint: position, orientation, linear momentum, angular momentum
enum { B_POS = 0, B_ORI = 3, B_LM = 7, B_AM = 10, B_COUNT = 13 };
// adjoint of the body state (orientation as a rotation-vector tangent)
enum { GB_POS = 0, GB_ORI = 3, GB_LM = 6, GB_AM = 9, GB_COUNT = 12 };

struct Plan {
  std::vector<int32_t> i;
  std::vector<double> d;
  int frames = 0;
  std::string error;  // non-empty: this problem has no re-rolled form (the flat tape path must be used)
};

inline int padTo(int n, int m) { return (n + m - 1) / m * m; }
// row stride (doubles) for shared-memory matrices read as DMMA fragments: == 4 (mod 16) keeps the 16 lanes of a
// half warp on distinct 8-byte banks for both the [row][k] and the transposed access (trx_rollout.cu)
inline int fragStride(int n) {
  int s = padTo(n, 8);
  while (s % 16 != 4) s++;
  return s;
}

inline Plan makePlan(const dexsyn::Model &m, const dexsyn::EnvInfo &env, int frames) {
  using namespace trxvm;
  Plan p;
  p.frames = frames;
  if (m.dropout) {
    p.error = "dropout masks are per-lane parameters of the flat tape only";
    return p;
  }
  if (m.wants_collision_pairs || m.friction_cones) {
    p.error = "collision and friction-cone penalties (collision_axes, collision_project_2) run on the flat tape only";
    return p;
  }
  const int n_all = (int)m.joints.size();
  // ---- which links does the objective look at?  (everything else is dead code on the reference tape too)
  std::vector<char> need(n_all, 0);
  auto mark = [&](int j) {
    while (j >= 0 && !need[j]) {
      need[j] = 1;
      j = m.joints[j].parent;
    }
  };
  mark(m.object_joint);
  const bool has_policy = m.policy_hidden >= 0;
  if (has_policy) {
    mark(m.palm_link);
    for (int e : m.end_effectors) mark(e);
  }
  // static = pose is a constant: fixed joints below fixed joints, hanging off the world
  std::vector<char> is_static(n_all, 0);
  for (int j = 0; j < n_all; j++) {
    const auto &jt = m.joints[j];
    if (jt.parent >= j) {
      p.error = "joint table is not in parent-first order";
      return p;
    }
    bool parent_static = jt.parent < 0 || is_static[jt.parent];
    is_static[j] = (jt.type == dexsyn::JOINT_FIXED && parent_static) ? 1 : 0;
  }
  // The moving part of the tree is cut into *segments*: maximal runs of joints in which every joint is the
  // only needed child of its predecessor.  One lane walks a segment from its first joint to its last with the
  // running pose in registers (no shared-memory round trip or warp barrier between the joints of a run);
  // segments that hang off a joint of another segment form the next *super-level*.
  std::vector<std::vector<int>> kids(n_all);
  std::vector<int> roots;
  for (int j = 0; j < n_all; j++) {
    if (!need[j] || is_static[j] || m.joints[j].type == dexsyn::JOINT_FLOATING) continue;
    int par = m.joints[j].parent;
    if (par >= 0 && !is_static[par]) {
      if (m.joints[par].type == dexsyn::JOINT_FLOATING) {
        p.error = "links hanging off the floating joint";
        return p;
      }
      kids[par].push_back(j);
    } else {
      roots.push_back(j);
    }
  }
  struct Seg {
    std::vector<int> joints;
    int super;
  };
  std::vector<Seg> segs;
  {
    std::vector<std::pair<int, int>> todo;  // (first joint, super-level)
    for (int r : roots) todo.push_back({r, 1});
    for (size_t t = 0; t < todo.size(); t++) {
      Seg sg;
      sg.super = todo[t].second;
      int j = todo[t].first;
      for (;;) {
        sg.joints.push_back(j);
        if (kids[j].size() == 1) {
          j = kids[j][0];
          continue;
        }
        for (int c : kids[j]) todo.push_back({c, sg.super + 1});
        break;
      }
      segs.push_back(sg);
    }
    std::stable_sort(segs.begin(), segs.end(), [](const Seg &x, const Seg &y) { return x.super < y.super; });
  }
  const int n_super = segs.empty() ? 0 : segs.back().super;
  // pose table order: statics, floating joints, then the segments (each contiguous)
  std::vector<int> order, index(n_all, -1);
  auto place = [&](int j) {
    index[j] = (int)order.size();
    order.push_back(j);
  };
  for (int j = 0; j < n_all; j++)
    if (need[j] && is_static[j]) place(j);
  const int n_static = (int)order.size();
  for (int j = 0; j < n_all; j++)
    if (need[j] && !is_static[j] && m.joints[j].type == dexsyn::JOINT_FLOATING) place(j);
  for (auto &sg : segs)
    for (int j : sg.joints) place(j);
  const int nj = (int)order.size();
  const int nq = (int)m.controlled.size();
  const int n_ee = has_policy ? (int)m.end_effectors.size() : 0;
  // only controlled joints that the objective can see matter; the policy still outputs all of them
  for (int c = 0; c < nq; c++)
    if (!need[m.controlled[c]]) {
      // a controlled joint outside every needed chain: its position still feeds the policy input
      mark(m.controlled[c]);
      p.error = "controlled joint outside the end-effector chains";  // not produced by dexsyn::makeModel
      return p;
    }

  p.i.assign(H_COUNT, 0);
  auto &I = p.i;
  auto &D = p.d;
  I[H_NJ] = nj;
  I[H_NSTATIC] = n_static;
  I[H_KSEG0] = nj;
  for (auto &sg : segs)
    for (int j : sg.joints) I[H_KSEG0] = std::min(I[H_KSEG0], index[j]);
  I[H_NSUPER] = n_super;
  I[H_NSEG] = (int)segs.size();
  I[H_NQ] = nq;
  I[H_NEE] = n_ee;
  I[H_OBJ] = index[m.object_joint];
  I[H_PALM] = has_policy ? index[m.palm_link] : 0;
  I[H_CDIMS] = m.contact_dims;
  const int n_in = nq + 12, n_out = nq + n_ee * m.contact_dims;
  int n_layers = 0, n_hid = 0;
  if (m.policy_hidden > 0) {
    n_layers = 2;
    n_hid = m.policy_hidden;
  } else if (m.policy_hidden == 0) {
    n_layers = 1;
  }
  I[H_NLAYERS] = n_layers;
  I[H_NIN] = n_in;
  I[H_NHID] = n_hid;
  I[H_NOUT] = has_policy ? n_out : 0;
  // parameter vector layout (dexsyn_assemble.h initLayers: weights then bias, layer by layer)
  if (n_layers == 2) {
    I[H_W1] = 0;
    I[H_B1] = n_in * n_hid;
    I[H_W2] = I[H_B1] + n_hid;
    I[H_B2] = I[H_W2] + n_hid * n_out;
    I[H_NX] = I[H_B2] + n_out;
  } else if (n_layers == 1) {
    I[H_W1] = 0;
    I[H_B1] = n_in * n_out;
    I[H_W2] = I[H_B2] = 0;
    I[H_NX] = I[H_B1] + n_out;
  }
  I[H_NSCALAR] = I[H_NX];

  // ---- integer tables
  I[H_I_SUPER] = (int)I.size();
  for (int sl = 1, sg = 0; sl <= n_super + 1; sl++) {
    while (sg < (int)segs.size() && segs[sg].super < sl) sg++;
    I.push_back(sg);
  }
  I[H_I_SEG] = (int)I.size();
  for (auto &sg : segs) {
    I.push_back(index[sg.joints.front()]);
    I.push_back((int)sg.joints.size());
  }
  // tracked links: end-effector links, then the object
  std::vector<int> track;
  auto trackOf = [&](int pose_index) {
    for (size_t t = 0; t < track.size(); t++)
      if (track[t] == pose_index) return (int)t;
    track.push_back(pose_index);
    return (int)track.size() - 1;
  };
  std::vector<int> ee_track(n_ee, 0);
  for (int e = 0; e < n_ee; e++) ee_track[e] = trackOf(index[m.end_effectors[e]]);
  const int obj_track = has_policy ? trackOf(index[m.object_joint]) : 0;
  const int n_track = (int)track.size();
  I[H_NTRACK] = n_track;
  // branch children of a joint: first joints of other segments that hang off it (the joint that continues
  // its own segment is simply the next pose-table entry)
  std::vector<std::vector<int>> children(nj);
  for (auto &sg : segs) {
    int par = m.joints[sg.joints.front()].parent;
    if (par >= 0 && !is_static[par]) children[index[par]].push_back(index[sg.joints.front()]);
  }
  I[H_I_JOINT] = (int)I.size();
  I.resize(I.size() + (size_t)nj * JI_STRIDE, 0);
  std::vector<int> child_lists;
  for (int k = 0; k < nj; k++) {
    const auto &jt = m.joints[order[k]];
    int *ji = &I[I[H_I_JOINT] + k * JI_STRIDE];
    ji[JI_PARENT] = jt.parent >= 0 ? index[jt.parent] : -1;
    ji[JI_TYPE] = jt.type;
    ji[JI_CTRL] = jt.type == dexsyn::JOINT_REVOLUTE ? jt.control : nq;  // n_q: a slot nobody reads
    ji[JI_CHILD0] = (int)child_lists.size();
    ji[JI_NCHILD] = (int)children[k].size();
    ji[JI_TRACK] = -1;
    for (int t = 0; t < n_track; t++)
      if (track[t] == k) ji[JI_TRACK] = t;
    for (int c : children[k]) child_lists.push_back(c);
  }
  {
    // child lists live behind the joint table; JI_CHILD0 becomes an absolute offset into I
    int base = (int)I.size();
    for (int c : child_lists) I.push_back(c);
    for (int k = 0; k < nj; k++) I[I[H_I_JOINT] + k * JI_STRIDE + JI_CHILD0] += base;
  }
  I[H_I_CTRL] = (int)I.size();
  for (int c = 0; c < nq; c++) {
    const auto &jt = m.joints[m.controlled[c]];
    I.push_back(index[m.controlled[c]]);
    I.push_back(jt.couple_to >= 0 ? m.joints[jt.couple_to].control : c);
    I.push_back(-1);
  }
  for (int c = 0; c < nq; c++) {
    const int src = I[I[H_I_CTRL] + c * CI_STRIDE + CI_SRC];
    if (src == c) continue;
    int &dst = I[I[H_I_CTRL] + src * CI_STRIDE + CI_DST];
    if (dst >= 0 || I[I[H_I_CTRL] + src * CI_STRIDE + CI_SRC] != src) {
      p.error = "joint coupling is not one-to-one";
      return p;
    }
    dst = c;
  }
  I[H_I_TRACK] = (int)I.size();
  for (int t : track) I.push_back(t);
  {
    // warp rounds: super-levels in order, segments first-fit into rounds of 32 lanes
    std::vector<int32_t> rounds, steps;
    int lane = 32, longest = 0;
    auto closeRound = [&]() {
      if (lane == 32 && rounds.empty()) return;
      int st = 0;
      while ((1 << st) < longest) st++;
      steps.push_back(st);
    };
    int cur_super = -1;
    for (auto &sg : segs) {
      const int n = (int)sg.joints.size();
      if (n > 32) {
        p.error = "a serial run of more than 32 joints";
        return p;
      }
      if (sg.super != cur_super || lane + n > 32) {
        if (!rounds.empty()) closeRound();
        rounds.resize(rounds.size() + 32 * RD_STRIDE, 0);
        for (int l = 0; l < 32; l++) rounds[rounds.size() - 32 * RD_STRIDE + l * RD_STRIDE + RD_K] = -1;
        lane = 0;
        longest = 0;
        cur_super = sg.super;
      }
      const int par = m.joints[sg.joints.front()].parent;
      for (int j = 0; j < n; j++, lane++) {
        int32_t *rd = &rounds[rounds.size() - 32 * RD_STRIDE + lane * RD_STRIDE];
        rd[RD_K] = index[sg.joints[j]];
        rd[RD_J] = j;
        rd[RD_LEN] = n;
        rd[RD_PAR] = par >= 0 ? index[par] : -1;
      }
      longest = std::max(longest, n);
    }
    if (!rounds.empty()) closeRound();
    for (auto &sg : segs)
      if (m.joints[sg.joints.front()].parent < 0) {
        p.error = "moving joint attached to the world";  // not produced by dexsyn::makeModel (the base link is fixed)
        return p;
      }
    while (I.size() % 4) I.push_back(0);
    I[H_NROUND] = (int)steps.size();
    I[H_I_ROUND] = (int)I.size();
    I.insert(I.end(), rounds.begin(), rounds.end());
    I[H_I_ROUNDSTEPS] = (int)I.size();
    I.insert(I.end(), steps.begin(), steps.end());
  }

  // ---- double tables
  D.clear();
  I[H_D_JOINT] = 0;
  D.resize((size_t)nj * JD_STRIDE, 0.0);
  for (int k = 0; k < nj; k++) {
    const auto &jt = m.joints[order[k]];
    double *jd = &D[(size_t)k * JD_STRIDE];
    for (int c = 0; c < 7; c++) jd[JD_ORIGIN + c] = jt.origin[c];
    for (int c = 0; c < 3; c++) jd[JD_AXIS + c] = jt.axis[c];
    Q4 r{jt.origin[3], jt.origin[4], jt.origin[5], jt.origin[6]};
    V3 ca = qrot(r, V3{jt.axis[0], jt.axis[1], jt.axis[2]});
    if (jt.type != dexsyn::JOINT_REVOLUTE) ca = V3{0, 0, 0};
    jd[JD_CAXIS + 0] = ca.x;
    jd[JD_CAXIS + 1] = ca.y;
    jd[JD_CAXIS + 2] = ca.z;
  }
  I[H_D_CTRL] = (int)D.size();
  for (int c = 0; c < nq; c++) {
    const auto &jt = m.joints[m.controlled[c]];
    D.push_back(jt.lower);
    D.push_back(jt.upper);
    D.push_back(jt.q0);
    D.push_back(jt.arm ? 3.0 : 20.0);  // dexenv_grasp.h:212-219
  }
  // shapes
  std::vector<double> planes;
  auto shapeRows = [&](const dexsyn::Shape &s) { return s.kind == 1 ? 9 : (int)s.planes.size(); };
  auto putShape = [&](const dexsyn::Shape &s, int link, int trk, std::vector<int32_t> &iv, std::vector<double> &dv) {
    iv.push_back(link);
    iv.push_back(s.kind);
    iv.push_back((int)s.planes.size());
    iv.push_back((int)planes.size() / 4);
    iv.push_back(shapeRows(s));
    iv.push_back(trk);
    iv.push_back(0);  // EI_ROW, filled below
    for (auto &pl : s.planes) {
      planes.push_back(pl.n[0]);
      planes.push_back(pl.n[1]);
      planes.push_back(pl.n[2]);
      planes.push_back(pl.d);
    }
    dv.push_back(s.center[0]);
    dv.push_back(s.center[1]);
    dv.push_back(s.center[2]);
    dv.push_back(s.radius);
    dv.push_back(s.length);
  };
  std::vector<int32_t> ee_i, obj_i;
  std::vector<double> ee_d, obj_d;
  for (int e = 0; e < n_ee; e++) putShape(m.ee_shapes[e], index[m.end_effectors[e]], ee_track[e], ee_i, ee_d);
  putShape(m.object_shape, index[m.object_joint], obj_track, obj_i, obj_d);
  // residual rows of one frame, in goal order (dexsyn_assemble.h rollout())
  int row = 2 * nq;
  I[H_ROW_ACT] = row;
  if (n_layers == 2) row += n_hid;
  I[H_ROW_CMD] = row;
  if (has_policy) row += n_out;
  I[H_ROW_EE0] = row;
  for (int e = 0; e < n_ee; e++) {
    ee_i[e * EI_STRIDE + EI_ROW] = row;
    row += 9 + obj_i[EI_NROWS] + ee_i[e * EI_STRIDE + EI_NROWS] + 18;
  }
  I[H_ROWS_FRAME] = row;
  I[H_I_EE] = (int)I.size();
  I.insert(I.end(), ee_i.begin(), ee_i.end());
  I[H_I_OBJSHAPE] = (int)I.size();
  I.insert(I.end(), obj_i.begin(), obj_i.end());
  I[H_D_EE] = (int)D.size();
  D.insert(D.end(), ee_d.begin(), ee_d.end());
  I[H_D_OBJSHAPE] = (int)D.size();
  D.insert(D.end(), obj_d.begin(), obj_d.end());
  I[H_D_PLANES] = (int)D.size();
  D.insert(D.end(), planes.begin(), planes.end());

  // constants.  Products of constants are folded at record time by the reference
  // (tractor/src/core/recorder.cpp:117-199), i.e. evaluated once in exactly this arithmetic.
  I[H_D_CONST] = (int)D.size();
  D.resize(D.size() + DC_COUNT, 0.0);
  double *dc = &D[I[H_D_CONST]];
  dc[DC_DT] = env.time_step;
  dc[DC_INV_DT] = 1.0 / env.time_step;
  {
    double mdt = m.body_mass * env.time_step;
    for (int c = 0; c < 3; c++) dc[DC_GRAV + c] = env.gravity[c] * mdt;
    dc[DC_DAMP_LIN] = std::exp(std::log(env.decay) * env.time_step);
    dc[DC_DAMP_ANG] = dc[DC_DAMP_LIN];
    dc[DC_INV_MASS] = 1.0 / m.body_mass;
  }
  for (int c = 0; c < 9; c++) dc[DC_IINV + c] = m.body_inverse_inertia[c];
  for (int c = 0; c < 3; c++) dc[DC_CENTER + c] = m.body_center[c];
  for (int c = 0; c < 7; c++) dc[DC_ANCHOR + c] = m.inverse_anchor[c];
  for (int c = 0; c < 7; c++) dc[DC_OBJ_ORIGIN + c] = m.joints[m.object_joint].origin[c];
  dc[DC_W_JLIMIT] = env.joint_limit_penalty;
  dc[DC_W_SHAPE] = env.shape_penalty;
  dc[DC_W_CPR] = env.contact_point_regularization;
  dc[DC_W_CFR] = env.contact_force_regularization;
  dc[DC_W_CDP] = env.contact_distance_penalty;
  dc[DC_W_CSP] = env.contact_slip_penalty;
  dc[DC_W_REG] = env.regularization;
  dc[DC_W_CREG] = env.command_regularization;
  for (int c = 0; c < 3; c++) dc[DC_GOAL + c] = env.move_goal[c];
  dc[DC_W_MOVE] = env.move_goal_weight;
  dc[DC_W_ORI] = env.orientation_goal_weight;
  {
    // object link pose of the initial FK: floating joint at the origin of its parent frame
    // (float_position = 0, float_orientation = identity): pmul(pmul(origin, ptrans(0)), porient(I))
    const auto &jt = m.joints[m.object_joint];
    if (jt.parent >= 0) {
      p.error = "floating joint must hang off the world";
      return p;
    }
    PoseV o{{jt.origin[0], jt.origin[1], jt.origin[2]}, {jt.origin[3], jt.origin[4], jt.origin[5], jt.origin[6]}};
    PoseV init = pmul(pmul(o, PoseV{{0, 0, 0}, {0, 0, 0, 1}}), PoseV{{0, 0, 0}, {0, 0, 0, 1}});
    double *q = dc + DC_OBJ_INIT;
    q[0] = init.t.x, q[1] = init.t.y, q[2] = init.t.z, q[3] = init.r.x, q[4] = init.r.y, q[5] = init.r.z, q[6] = init.r.w;
  }
  // constant poses of the static links (composition of fixed origins from the world down)
  I[H_D_STATIC] = (int)D.size();
  {
    std::vector<PoseV> sp(n_static);
    for (int k = 0; k < n_static; k++) {
      const auto &jt = m.joints[order[k]];
      PoseV o{{jt.origin[0], jt.origin[1], jt.origin[2]}, {jt.origin[3], jt.origin[4], jt.origin[5], jt.origin[6]}};
      sp[k] = jt.parent >= 0 ? pmul(sp[index[jt.parent]], o) : o;
      D.push_back(sp[k].t.x);
      D.push_back(sp[k].t.y);
      D.push_back(sp[k].t.z);
      D.push_back(sp[k].r.x);
      D.push_back(sp[k].r.y);
      D.push_back(sp[k].r.z);
      D.push_back(sp[k].r.w);
    }
  }

  // ---- per-rollout region
  int off = 0;
  auto take = [&](int key, int n) {
    I[key] = off;
    off += n;
  };
  // carried state first, in checkpoint order: a checkpoint is one contiguous copy of it
  take(R_Q, nq);
  take(R_CMD, nq);
  take(R_BODY, B_COUNT);
  take(R_PREV, n_track * 7);
  // fields that bulk copies restore from a checkpoint (trx_rollout.cu) start on 16-byte boundaries and own an
  // even number of doubles, so that a copy rounded up to 16 bytes stays inside its field
  auto takeEven = [&](int key, int n) {
    off = padTo(off, 2);
    I[key] = off;
    off += padTo(n, 2);
  };
  takeEven(R_POSE, nj * 7);
  take(R_LQ, nj * 4);
  take(R_GPOSE, nj * 6);
  take(R_GC, nj * 6);
  takeEven(R_VEL, std::max(nq, 1));
  take(R_GQ, nq + 1);
  take(R_GCMD, nq);
  take(R_GBODY, GB_COUNT);
  takeEven(R_CG, 3);
  take(R_GCG, 3);
  take(R_FIRST, 7);
  take(R_GFIRST, 6);
  take(R_GPREV, n_track * 6);
  takeEven(R_EE, n_ee * EE_STRIDE);
  take(R_GOBJ, n_ee * 6);
  take(R_GPREVOBJ, n_ee * 6);
  take(R_BODY0, B_COUNT);
  // region stride == 2 (mod 16) doubles: the kernel's lanes hold the same item of 8 rollouts (trx_rollout.cu), so
  // a half warp reads one offset of 8 regions plus the next item's -- distinct 8-byte banks
  I[H_REGION] = padTo(off, 2);
  while (I[H_REGION] % 16 != 2) I[H_REGION] += 2;
  // checkpoint: joint positions, commands, body state, tracked previous poses
  I[H_CKPT] = nq + nq + B_COUNT + n_track * 7;
  return p;
}

}  // namespace rollout
