// tape_builder.h -- the product's own tape recorder.
//
// Backend for the objective assembly (csrc/assembly/dexsyn_assemble.h) that writes a
// tractor-style tape directly in the neutral program form, without the reference
// library: what Var<> + Recorder do in the reference
// (tractor/include/tractor/core/var.h:12-81, core/recorder.h:59-132,
// src/core/recorder.cpp:479-612) reduced to what a GPU engine needs --
//   * every operation appends [operator, argument addresses] and gets a fresh
//     write-once slot for each output (recorder.cpp:485-496,552)
//   * literals are constants with their bytes in the constant blob (recorder.h:88-104)
//   * weights are scalar-typed input ports, random draws are lane-typed parameter
//     ports, goals are output ports (recorder.h:66-132)
//   * finish() runs constant folding and dead-code removal (tape_passes.h)
// Handles are plain addresses, so no Var-copy `move` instructions are ever recorded
// (the reference removes those afterwards with skipMoves, recorder.cpp:201-250).
// Operator names are the reference's, so the resulting program runs on any tractor
// engine and can be compared with a reference recording of the same assembly.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "tape_gradients.h"

namespace trxb {

struct Handle {
  uint64_t addr = ~0ull;
};

class TapeBuilder {
 public:
  struct S : Handle {};
  struct W : Handle {};
  struct V3 : Handle {};
  struct Q : Handle {};
  struct P : Handle {};
  struct M3 : Handle {};

  static constexpr bool kHasCollisionOps = true;
  trxfile::Program prog;
  std::vector<uint64_t> frame_begin_inst;  // instruction count at frameBegin(t), before the passes
  // record a dense layer as one block instruction (x_dense, vm_ops.def) instead of the reference's
  // per-weight batch + dot4add expansion
  bool fuse_dense = false;

 private:
  Alloc _alloc;
  std::map<std::string, uint32_t> _index;
  static constexpr uint32_t L = 8;

  uint64_t allocT(uint32_t comps) { return _alloc.alloc(comps * 8 * L, L); }
  uint64_t allocS() { return _alloc.alloc(8, 1); }

  template <class H> H constantT(const std::vector<double> &comps) {
    H h;
    h.addr = allocT((uint32_t)comps.size());
    trxfile::Port c;
    c.address = h.addr;
    c.offset = prog.const_data.size();
    c.size = (uint32_t)comps.size() * 8 * L;
    c.lanes = L;
    prog.constants.push_back(c);
    for (double v : comps)
      for (uint32_t l = 0; l < L; l++) {
        const uint8_t *b = (const uint8_t *)&v;
        prog.const_data.insert(prog.const_data.end(), b, b + 8);
      }
    return h;
  }

  // appends one instruction; outputs get fresh slots sized from the operator table
  void emit(const std::string &name, const std::vector<uint64_t> &inputs, std::vector<uint64_t> &outputs) {
    uint32_t id = internOp(prog, _index, name);
    const auto &op = prog.ops[id];
    prog.code.push_back(id);
    size_t in_i = 0;
    outputs.clear();
    for (auto &a : op.args) {
      if (a.is_input) {
        if (in_i >= inputs.size()) throw std::runtime_error("too few inputs for " + name);
        if (inputs[in_i] == ~0ull) throw std::runtime_error("uninitialised handle passed to " + name);
        prog.code.push_back(inputs[in_i++]);
      } else {
        uint64_t addr = _alloc.alloc(a.size, a.lanes);
        outputs.push_back(addr);
        prog.code.push_back(addr);
      }
    }
    if (in_i != inputs.size()) throw std::runtime_error("too many inputs for " + name);
    prog.n_instructions++;
  }
  template <class H> H op1(const std::string &name, const std::vector<uint64_t> &inputs) {
    std::vector<uint64_t> outs;
    emit(name, inputs, outs);
    if (outs.size() != 1) throw std::runtime_error("expected one output from " + name);
    H h;
    h.addr = outs[0];
    return h;
  }
  void outputPort(uint64_t addr, uint32_t size, uint32_t lanes) {
    trxfile::Port p;
    p.address = addr;
    p.size = size;
    p.lanes = (uint8_t)lanes;
    p.offset = prog.outputs.empty() ? 0 : prog.outputs.back().offset + prog.outputs.back().size;
    prog.goals.push_back({(uint64_t)prog.outputs.size(), 0});
    prog.outputs.push_back(p);
  }

 public:
  TapeBuilder() { prog.lane_width = L; }

  // ---- constants
  S cs(double v) { return constantT<S>({v}); }
  W cw(double v) {
    W h;
    h.addr = allocS();
    trxfile::Port c;
    c.address = h.addr;
    c.offset = prog.const_data.size();
    c.size = 8;
    c.lanes = 1;
    prog.constants.push_back(c);
    const uint8_t *b = (const uint8_t *)&v;
    prog.const_data.insert(prog.const_data.end(), b, b + 8);
    return h;
  }
  V3 cv(double x, double y, double z) { return constantT<V3>({x, y, z}); }
  Q cq(double x, double y, double z, double w) { return constantT<Q>({x, y, z, w}); }
  P cp(const double *o) { return constantT<P>({o[0], o[1], o[2], o[3], o[4], o[5], o[6]}); }
  M3 cm(const double *m) { return constantT<M3>(std::vector<double>(m, m + 9)); }

  // ---- ports
  W weight(double) {
    W h;
    h.addr = allocS();
    trxfile::Port p;
    p.address = h.addr;
    p.size = 8;
    p.lanes = 1;
    p.offset = prog.inputs.empty() ? 0 : prog.inputs.back().offset + prog.inputs.back().size;
    prog.inputs.push_back(p);
    return h;
  }
  S param() {
    S h;
    h.addr = allocT(1);
    trxfile::Port p;
    p.address = h.addr;
    p.size = 8 * L;
    p.lanes = L;
    p.offset = prog.parameters.empty() ? 0 : prog.parameters.back().offset + prog.parameters.back().size;
    prog.parameters.push_back(p);
    return h;
  }
  void goal(const S &v) { outputPort(v.addr, 8 * L, L); }
  void goalV(const V3 &v) { outputPort(v.addr, 3 * 8 * L, L); }
  void goalW(const W &v) { outputPort(v.addr, 8, 1); }

  void frameBegin(int) { frame_begin_inst.push_back(prog.n_instructions); }
  void frameEnd(int) {}

  // ---- scalar ops
  S add(const S &a, const S &b) { return op1<S>("add_8d", {a.addr, b.addr}); }
  S sub(const S &a, const S &b) { return op1<S>("sub_8d", {a.addr, b.addr}); }
  S mul(const S &a, const S &b) { return op1<S>("mul_8d", {a.addr, b.addr}); }
  S div(const S &a, const S &b) { return op1<S>("div_8d", {a.addr, b.addr}); }
  S neg(const S &a) { return op1<S>("minus_8d", {a.addr}); }
  S tanh(const S &a) { return op1<S>("tanh_8d", {a.addr}); }
  S relu(const S &a) { return op1<S>("relu_8d", {a.addr}); }
  S exp(const S &a) { return op1<S>("exp_8d", {a.addr}); }
  S log(const S &a) { return op1<S>("log_8d", {a.addr}); }
  W wmul(const W &a, const W &b) { return op1<W>("mul_d", {a.addr, b.addr}); }

  // ---- vectors
  V3 vadd(const V3 &a, const V3 &b) { return op1<V3>("add_fg_vec3_8d", {a.addr, b.addr}); }
  V3 vsub(const V3 &a, const V3 &b) { return op1<V3>("sub_fg_vec3_8d", {a.addr, b.addr}); }
  V3 vneg(const V3 &a) { return op1<V3>("minus_fg_vec3_8d", {a.addr}); }
  V3 vmuls(const V3 &a, const S &b) { return op1<V3>("mul_fg_vec3_s_8d", {a.addr, b.addr}); }
  S dot(const V3 &a, const V3 &b) { return op1<S>("dot_fg_vec3_8d", {a.addr, b.addr}); }
  V3 cross(const V3 &a, const V3 &b) { return op1<V3>("cross_fg_vec3_8d", {a.addr, b.addr}); }
  V3 pack(const S &x, const S &y, const S &z) { return op1<V3>("fg_vec3_pack_8d", {x.addr, y.addr, z.addr}); }
  void unpack(const V3 &v, S &x, S &y, S &z) {
    std::vector<uint64_t> outs;
    emit("fg_vec3_unpack_8d", {v.addr}, outs);
    x.addr = outs[0];
    y.addr = outs[1];
    z.addr = outs[2];
  }

  // ---- orientations
  V3 qmulv(const Q &q, const V3 &v) { return op1<V3>("mul_fg_quat_vec3_8d", {q.addr, v.addr}); }
  Q qmul(const Q &a, const Q &b) { return op1<Q>("mul_fg_quat_8d", {a.addr, b.addr}); }
  Q qinv(const Q &a) { return op1<Q>("fg_quat_inverse_8d", {a.addr}); }
  Q qaxis(const S &angle, const V3 &axis) { return op1<Q>("fg_angle_axis_quat_8d", {angle.addr, axis.addr}); }
  Q qaddv(const Q &q, const V3 &v) { return op1<Q>("add_fg_quat_vec3_8d", {q.addr, v.addr}); }
  V3 qresidual(const Q &q) { return op1<V3>("fg_quat_residual_8d", {q.addr}); }

  // ---- poses
  P pmul(const P &a, const P &b) { return op1<P>("mul_fg_pose_8d", {a.addr, b.addr}); }
  V3 pmulv(const P &a, const V3 &v) { return op1<V3>("mul_fg_pose_vec3_8d", {a.addr, v.addr}); }
  P prevolute(const P &parent, const S &angle, const V3 &axis) {
    return op1<P>("fg_pose_angle_axis_pose_8d", {parent.addr, angle.addr, axis.addr});
  }
  P ptrans(const V3 &t) { return op1<P>("fg_translation_pose_8d", {t.addr}); }
  P porient(const Q &q) { return op1<P>("fg_orientation_pose_8d", {q.addr}); }
  V3 ptranslation(const P &p) { return op1<V3>("fg_pose_translation_8d", {p.addr}); }
  Q porientation(const P &p) { return op1<Q>("fg_pose_orientation_8d", {p.addr}); }
  V3 mmulv(const M3 &m, const V3 &v) { return op1<V3>("mul_fg_mat3_vec3_8d", {m.addr, v.addr}); }

  // ---- collision operators (dexopt/src/ops.h:208-252,288-375).  The uint64 constant is the shape / pair handle;
  // every instruction gets its own constant slot so that no gradient slot of a handle has two writers.
  uint64_t cu(uint64_t v) {
    uint64_t addr = allocS();
    trxfile::Port c;
    c.address = addr;
    c.offset = prog.const_data.size();
    c.size = 8;
    c.lanes = 1;
    c.elem = trxfile::ELEM_U64;
    prog.constants.push_back(c);
    const uint8_t *b = (const uint8_t *)&v;
    prog.const_data.insert(prog.const_data.end(), b, b + 8);
    return addr;
  }
  void collisionProject(const V3 &point, uint64_t shape, V3 &normal, S &distance) {
    std::vector<uint64_t> outs;
    emit("collision_project_2_8d", {point.addr, cu(shape)}, outs);
    normal.addr = outs[0];
    distance.addr = outs[1];
  }
  void collisionAxes(const P &pose_a, const P &pose_b, uint64_t pair, V3 &point_a, V3 &point_b, V3 &axis, V3 &local_a,
                     V3 &local_b) {
    std::vector<uint64_t> outs;
    emit("collision_axes_8d", {pose_a.addr, pose_b.addr, cu(pair)}, outs);
    point_a.addr = outs[0];
    point_b.addr = outs[1];
    axis.addr = outs[2];
    local_a.addr = outs[3];
    local_b.addr = outs[4];
  }

  // ---- dense layer, expanded exactly like the reference records it
  // (dexopt/src/tensor.h:134-172 and dexopt/src/neural.h:176-181)
  std::vector<S> dense(const std::vector<S> &a, const std::vector<W> &wts, const std::vector<W> &bias, int n_in,
                       int n_out) {
    if (fuse_dense) return denseFused(a, wts, bias, n_in, n_out);
    std::vector<S> r;
    S zero = cs(0.0);
    for (int col = 0; col < n_out; col++) {
      S acc = zero;
      int row = 0;
      for (; row + 3 < n_in; row += 4) {
        S w[4];
        for (int i = 0; i < 4; i++) w[i] = op1<S>("batch_8d", {wts[(row + i) * n_out + col].addr});
        acc = op1<S>("dot4add_8d", {a[row].addr, a[row + 1].addr, a[row + 2].addr, a[row + 3].addr, w[0].addr,
                                    w[1].addr, w[2].addr, w[3].addr, acc.addr});
      }
      for (; row < n_in; row++) {
        S w = op1<S>("batch_8d", {wts[row * n_out + col].addr});
        acc = op1<S>("madd_8d", {a[row].addr, w.addr, acc.addr});
      }
      r.push_back(acc);
    }
    for (int i = 0; i < n_out; i++) {
      S bv = op1<S>("batch_8d", {bias[i].addr});
      r[i] = add(r[i], bv);
    }
    return r;
  }

  // One instruction per layer: gather the activations into a contiguous block (plain moves), then
  // x_dense on [x block, weight block, bias block] -> y block.  The weights of a layer are consecutive
  // scalar input ports by construction (Assembler::initLayers).
  std::vector<S> denseFused(const std::vector<S> &a, const std::vector<W> &wts, const std::vector<W> &bias, int n_in,
                            int n_out) {
    for (size_t k = 1; k < wts.size(); k++)
      if (wts[k].addr != wts[0].addr + 8 * k) throw std::runtime_error("dense layer weights are not contiguous");
    for (size_t k = 1; k < bias.size(); k++)
      if (bias[k].addr != bias[0].addr + 8 * k) throw std::runtime_error("dense layer biases are not contiguous");
    while (_alloc.top % 64) _alloc.top++;
    uint64_t xb = _alloc.alloc((uint32_t)n_in * 64, L);
    uint32_t mv = internOp(prog, _index, "move_8d");
    for (int i = 0; i < n_in; i++) {
      prog.code.push_back(mv);
      prog.code.push_back(a[i].addr);
      prog.code.push_back(xb + 64 * (uint64_t)i);
      prog.n_instructions++;
    }
    uint64_t yb = _alloc.alloc((uint32_t)n_out * 64, L);
    uint32_t id = internBlockOp(prog, _index, "x_dense_8d",
                                {(uint32_t)n_in * 64, (uint32_t)(n_in * n_out) * 8, (uint32_t)n_out * 8, (uint32_t)n_out * 64});
    prog.code.push_back(id);
    prog.code.push_back(xb);
    prog.code.push_back(wts[0].addr);
    prog.code.push_back(bias[0].addr);
    prog.code.push_back(yb);
    prog.n_instructions++;
    std::vector<S> r(n_out);
    for (int j = 0; j < n_out; j++) r[j].addr = yb + 64 * (uint64_t)j;
    return r;
  }

  // recorder.cpp:586-611
  void finish() {
    prog.memory_size = _alloc.top;
    foldConstants(prog);
    removeUnused(prog);
  }
};

}  // namespace trxb
