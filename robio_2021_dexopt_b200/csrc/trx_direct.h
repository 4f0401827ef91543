// trx_direct.h -- the adjoint reads forward values in place (objective lowering).
//
// The reference splits reverse mode into x_prep and x_bprop (tractor/src/core/gradients.cpp:57-118): every
// `prepare_*` rule saves what its `reverse_*` rule needs into a private slot of the prep memory, and for many
// operators that is a plain copy of an operand (mul saves {a, b}; mul_fg_vec3_s, dot_fg_vec3, cross_fg_vec3,
// mul_fg_quat_vec3, ... save a and b; gate saves b; fg_pose_translate saves the orientation of a).
// On one Memory the copy is redundant: prog has finished before bprop starts and neither prep nor bprop writes
// forward slots, so the reverse rule can read the operand where prog left it.  Inside an objective (the three
// tapes are known together and only ports cross programs) this pass
//   * redirects every bprop argument that names such a copy to the forward slot it was copied from,
//   * replaces reverse_mul (whose saved state is ONE two-component slot) by reverse_x_muld, the same rule
//     with the two operands as separate arguments (vm_ops.def),
//   * drops the prepare instructions all of whose outputs became unused.
// Values are untouched: every reverse rule sees bit-identical operands.  What disappears is the copy traffic
// (a DRAM read and a DRAM write per saved operand) and about three quarters of the prep launch.
#pragma once
#include <algorithm>
#include <cstdint>
#include <initializer_list>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/trx_file.h"
#include "trx_ops.h"

namespace trx {

struct DirectStats {
  uint64_t n_prepare_dropped = 0, n_arguments_redirected = 0, n_mul_rewritten = 0;
};

inline void directPrepare(trxfile::Program &prep, trxfile::Program &bprop, DirectStats &st) {
  // saved slot (output k_out of the prepare rule) = components [comp_offset, comp_offset + comps) of input k_in
  struct Copy {
    int k_out, k_in, comp_offset;
  };
  static const std::map<std::string, std::vector<Copy>> kCopies = {
      {"mul_fg_vec3_s", {{3, 0, 0}, {4, 1, 0}}},     {"mul_fg_s_vec3", {{3, 0, 0}, {4, 1, 0}}},
      {"dot_fg_vec3", {{3, 0, 0}, {4, 1, 0}}},       {"cross_fg_vec3", {{3, 0, 0}, {4, 1, 0}}},
      {"mul_fg_quat_vec3", {{3, 0, 0}, {4, 1, 0}}},  {"mul_fg_quat", {{3, 0, 0}}},
      {"fg_angle_axis_quat", {{3, 0, 0}, {4, 1, 0}}}, {"mul_fg_twist_s", {{3, 0, 0}, {4, 1, 0}}},
      {"fg_pose_translate", {{3, 0, 3}}},            {"gate", {{3, 1, 0}}},
      {"gate_fg_pose_gate", {{3, 1, 0}}},
  };
  struct Inst {
    uint32_t op;
    uint64_t at;  // first argument word
    uint32_t n;
  };
  auto decode = [](const trxfile::Program &p) {
    std::vector<Inst> v;
    v.reserve(p.n_instructions);
    for (uint64_t pc = 0; pc < p.code.size();) {
      uint64_t oi = p.code[pc];
      if (oi >= p.ops.size()) throw std::runtime_error("invalid op");
      uint32_t n = (uint32_t)p.ops[oi].args.size();
      v.push_back({(uint32_t)oi, pc + 1, n});
      pc += 1 + n;
    }
    return v;
  };
  // "prepare_<base>_<8d|d>" -> base (empty when the name has another form)
  auto prepareBase = [](const std::string &name) -> std::string {
    if (name.rfind("prepare_", 0) != 0) return "";
    size_t e = name.rfind('_');
    if (e == std::string::npos || e <= 8) return "";
    return name.substr(8, e - 8);
  };
  const auto pi = decode(prep), bi = decode(bprop);
  // slots written by prep / bprop (an operand that either of them overwrites is not safe to read late),
  // write counts of prep outputs, and port slots (they are visible outside the objective's programs)
  std::unordered_map<uint64_t, uint32_t> prep_writes;
  std::unordered_set<uint64_t> written, ports;
  for (auto &in : pi)
    for (uint32_t k = 0; k < in.n; k++)
      if (!prep.ops[in.op].args[k].is_input) {
        prep_writes[prep.code[in.at + k]]++;
        written.insert(prep.code[in.at + k]);
      }
  std::unordered_set<uint64_t> bprop_written;
  for (auto &in : bi)
    for (uint32_t k = 0; k < in.n; k++)
      if (!bprop.ops[in.op].args[k].is_input) {
        written.insert(bprop.code[in.at + k]);
        bprop_written.insert(bprop.code[in.at + k]);
      }
  for (auto *prog : {&prep, &bprop})
    for (auto *pp : {&prog->inputs, &prog->parameters, &prog->outputs, &prog->constants})
      for (auto &p : *pp)
        for (uint64_t a = p.address; a < p.address + p.size; a += 8) ports.insert(a);
  auto safeSaved = [&](uint64_t s) { return prep_writes[s] == 1 && !bprop_written.count(s) && !ports.count(s); };

  std::unordered_map<uint64_t, uint64_t> alias;                      // saved slot -> forward address
  std::unordered_map<uint64_t, std::pair<uint64_t, uint64_t>> mul2;  // saved {a, b} slot of mul -> (a, b)
  std::vector<std::vector<uint8_t>> out_aliased(pi.size());
  for (size_t i = 0; i < pi.size(); i++) {
    const Inst &in = pi[i];
    const auto &op = prep.ops[in.op];
    out_aliased[i].assign(in.n, 0);
    const std::string base = prepareBase(op.name);
    if (base == "mul" && in.n == 4) {
      const uint64_t a = prep.code[in.at], b = prep.code[in.at + 1], s = prep.code[in.at + 3];
      if (op.args[3].size == 2 * op.args[0].size && safeSaved(s) && !written.count(a) && !written.count(b)) {
        mul2[s] = {a, b};
        out_aliased[i][3] = 1;
      }
      continue;
    }
    auto it = kCopies.find(base);
    if (it == kCopies.end()) continue;
    for (const Copy &c : it->second) {
      if ((uint32_t)c.k_out >= in.n || (uint32_t)c.k_in >= in.n) throw std::runtime_error("prepare rule layout changed: " + op.name);
      const auto &ao = op.args[c.k_out], &ai = op.args[c.k_in];
      if (ao.is_input || !ai.is_input || ao.lanes != ai.lanes) throw std::runtime_error("prepare rule layout changed: " + op.name);
      const uint64_t row = 8ull * ai.lanes;  // bytes per component
      if (c.comp_offset * row + ao.size > ai.size) throw std::runtime_error("prepare rule layout changed: " + op.name);
      const uint64_t s = prep.code[in.at + c.k_out], from = prep.code[in.at + c.k_in];
      if (!safeSaved(s) || written.count(from)) continue;
      alias[s] = from + c.comp_offset * row;
      out_aliased[i][c.k_out] = 1;
    }
  }
  if (alias.empty() && mul2.empty()) return;

  // ---- bprop: redirect arguments, rewrite reverse_mul
  std::vector<uint64_t> code;
  code.reserve(bprop.code.size());
  std::map<std::string, uint32_t> muld_ops;  // lane suffix -> op index of reverse_x_muld
  for (auto &in : bi) {
    const auto &op = bprop.ops[in.op];
    uint16_t opcode;
    uint32_t lw;
    const bool known = findOp(op.name, opcode, lw);
    if (known && opcode == OP_R_MUL && in.n == 4) {
      auto m = mul2.find(bprop.code[in.at]);
      if (m != mul2.end()) {
        const std::string suffix = lw == 8 ? "_8d" : "_d";
        auto it = muld_ops.find(suffix);
        if (it == muld_ops.end()) {
          trxfile::OpDesc d;
          d.name = "reverse_x_muld" + suffix;
          trxfile::ArgDesc operand = op.args[1];  // one scalar of the tape's type
          operand.is_input = 1;
          d.args = {operand, operand, op.args[1], op.args[2], op.args[3]};
          bprop.ops.push_back(d);
          it = muld_ops.emplace(suffix, (uint32_t)bprop.ops.size() - 1).first;
        }
        code.push_back(it->second);
        code.push_back(m->second.first);
        code.push_back(m->second.second);
        for (uint32_t k = 1; k < 4; k++) code.push_back(bprop.code[in.at + k]);
        st.n_mul_rewritten++;
        continue;
      }
    }
    code.push_back(in.op);
    for (uint32_t k = 0; k < in.n; k++) {
      uint64_t a = bprop.code[in.at + k];
      if (bprop.ops[in.op].args[k].is_input) {
        auto al = alias.find(a);
        if (al != alias.end()) {
          a = al->second;
          st.n_arguments_redirected++;
        } else if (mul2.count(a)) {
          // a reader of the {a, b} slot other than reverse_mul: keep the copy alive
          throw std::runtime_error("saved state of mul read by " + bprop.ops[in.op].name);
        }
      }
      code.push_back(a);
    }
  }
  bprop.code.swap(code);

  // ---- prep: drop the prepare instructions whose outputs are all unused now
  std::vector<uint64_t> pcode;
  pcode.reserve(prep.code.size());
  uint64_t kept = 0;
  for (size_t i = 0; i < pi.size(); i++) {
    const Inst &in = pi[i];
    bool all = false;
    bool any_out = false;
    for (uint32_t k = 0; k < in.n; k++)
      if (!prep.ops[in.op].args[k].is_input) {
        if (!any_out) all = true;
        any_out = true;
        all = all && out_aliased[i][k];
      }
    if (any_out && all) {
      st.n_prepare_dropped++;
      continue;
    }
    // (an instruction with only some outputs redirected still runs; its readers use whichever copy the
    // redirect map names -- the values are equal)
    pcode.push_back(in.op);
    for (uint32_t k = 0; k < in.n; k++) pcode.push_back(prep.code[in.at + k]);
    kept++;
  }
  prep.code.swap(pcode);
  prep.n_instructions = kept;
}

// True when no instruction output and no input / parameter port of the given programs overlaps a
// constant slot of any of them: an objective that runs exactly these programs on one Memory may then upload
// the constants once instead of with every execute (the reference copies them every time, simple.cpp:27-30).
inline bool constantsAreNeverWritten(std::initializer_list<const trxfile::Program *> programs) {
  std::vector<std::pair<uint64_t, uint64_t>> ranges;
  for (auto *p : programs)
    for (auto &c : p->constants) ranges.push_back({c.address, c.address + c.size});
  if (ranges.empty()) return true;
  std::sort(ranges.begin(), ranges.end());
  size_t w = 0;
  for (size_t i = 0; i < ranges.size(); i++) {
    if (w && ranges[i].first <= ranges[w - 1].second) ranges[w - 1].second = std::max(ranges[w - 1].second, ranges[i].second);
    else ranges[w++] = ranges[i];
  }
  ranges.resize(w);
  auto overlaps = [&](uint64_t a, uint64_t size) {
    if (!size) return false;
    auto it = std::lower_bound(ranges.begin(), ranges.end(), std::make_pair(a + size, (uint64_t)0));
    if (it == ranges.begin()) return false;
    --it;
    return it->second > a;
  };
  for (auto *p : programs) {
    // inputs and parameters are written by the port scatter (and bprop's inputs by the adjoint seeds);
    // outputs are only read
    for (auto *ports : {&p->inputs, &p->parameters})
      for (auto &q : *ports)
        if (overlaps(q.address, q.size)) return false;
    for (uint64_t pc = 0; pc < p->code.size();) {
      const uint64_t oi = p->code[pc];
      if (oi >= p->ops.size()) return false;
      const auto &op = p->ops[oi];
      for (size_t k = 0; k < op.args.size(); k++)
        if (!op.args[k].is_input && overlaps(p->code[pc + 1 + k], op.args[k].size)) return false;
      pc += 1 + op.args.size();
    }
  }
  return true;
}

}  // namespace trx
