// tape_passes.h -- tape clean-up passes of the product-side recorder.
//
// The reference runs four passes when a recording finishes
// (tractor/src/core/recorder.cpp:589-611): constant pre-computation (:117-199),
// move skipping (:201-250), unused-instruction removal (:63-115) and
// unused-constant removal (:36-61).  The product's tape builder never emits
// Var-copy moves, so three passes remain; they are restated here on the
// neutral program form (include/trx_file.h).  Constant folding evaluates the
// foldable instructions once on the host with the VM's own operator bodies
// (vm_exec.cuh built as host code) -- a compile-time use, exactly like the
// reference evaluating them at record time (recorder.cpp:162).
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../../include/trx_file.h"
#include "../trx_ops.h"
#include "../vm_exec.cuh"

namespace trxb {

struct DecodedInst {
  uint32_t op;       // index into Program::ops
  uint64_t code_at;  // index of first argument in Program::code
  uint32_t n_args;
};

inline std::vector<DecodedInst> decode(const trxfile::Program &p) {
  std::vector<DecodedInst> out;
  out.reserve(p.n_instructions);
  uint64_t pc = 0;
  while (pc < p.code.size()) {
    uint64_t oi = p.code[pc];
    if (oi >= p.ops.size()) throw std::runtime_error("invalid op");
    uint32_t na = (uint32_t)p.ops[oi].args.size();
    out.push_back({(uint32_t)oi, pc + 1, na});
    pc += 1 + na;
  }
  return out;
}

// interns an operator name into the program's operator table, filling the argument layout from
// the device operator vocabulary
inline uint32_t internOp(trxfile::Program &p, std::map<std::string, uint32_t> &index, const std::string &name) {
  auto it = index.find(name);
  if (it != index.end()) return it->second;
  uint16_t opcode;
  uint32_t lw;
  if (!trx::findOp(name, opcode, lw)) throw std::runtime_error("operator not found: " + name);
  const trx::OpSig &sig = trx::opTable()[opcode];
  trxfile::OpDesc d;
  d.name = name;
  for (auto &a : sig.args) {
    trxfile::ArgDesc ad;
    uint32_t lanes = (a.kind == trx::ARG_T) ? lw : 1;
    ad.size = a.comps * 8 * lanes;
    ad.is_input = a.is_input;
    ad.lanes = (uint8_t)lanes;
    ad.elem = (a.kind == trx::ARG_U) ? trxfile::ELEM_U64 : trxfile::ELEM_F64;
    d.args.push_back(ad);
  }
  uint32_t id = (uint32_t)p.ops.size();
  p.ops.push_back(d);
  index[name] = id;
  return id;
}

// block operators (vm_ops.def, "product extensions") carry their block sizes in the operator table
// entry, so every shape is its own entry; `sizes` = bytes per argument
inline uint32_t internBlockOp(trxfile::Program &p, std::map<std::string, uint32_t> &index, const std::string &name,
                              const std::vector<uint32_t> &sizes) {
  std::string key = name;
  for (auto s : sizes) key += "#" + std::to_string(s);
  auto it = index.find(key);
  if (it != index.end()) return it->second;
  uint16_t opcode;
  uint32_t lw;
  if (!trx::findOp(name, opcode, lw)) throw std::runtime_error("operator not found: " + name);
  const trx::OpSig &sig = trx::opTable()[opcode];
  if (sig.args.size() != sizes.size()) throw std::runtime_error("block operator arity mismatch: " + name);
  trxfile::OpDesc d;
  d.name = name;
  for (size_t k = 0; k < sizes.size(); k++) {
    trxfile::ArgDesc ad;
    ad.size = sizes[k];
    ad.is_input = sig.args[k].is_input;
    ad.lanes = (sig.args[k].kind == trx::ARG_TB) ? 8 : 1;
    ad.elem = trxfile::ELEM_F64;
    d.args.push_back(ad);
  }
  uint32_t id = (uint32_t)p.ops.size();
  p.ops.push_back(d);
  index[key] = id;
  return id;
}

struct OpMeta {
  uint16_t opcode = 0;
  uint32_t lane_width = 1;
  const trx::OpSig *sig = nullptr;
};
inline std::vector<OpMeta> opMeta(const trxfile::Program &p) {
  std::vector<OpMeta> m(p.ops.size());
  for (size_t i = 0; i < p.ops.size(); i++) {
    if (!trx::findOp(p.ops[i].name, m[i].opcode, m[i].lane_width))
      throw std::runtime_error("operator not found: " + p.ops[i].name);
    m[i].sig = &trx::opTable()[m[i].opcode];
  }
  return m;
}
// calls f(address) for the argument's start address, or for every element of a block argument
template <class F>
inline void forEachKey(const trxfile::Program &p, const std::vector<OpMeta> &meta, const DecodedInst &in, uint32_t k, F f) {
  const trx::ArgSig &a = meta[in.op].sig->args[k];
  uint64_t addr = p.code[in.code_at + k];
  if (!trx::isBlock(a.kind)) {
    f(addr);
    return;
  }
  uint32_t eb = trx::blockElemBytes(a.kind, 8), size = p.ops[in.op].args[k].size;
  for (uint32_t off = 0; off < size; off += eb) f(addr + off);
}

inline std::map<std::string, uint32_t> opIndex(const trxfile::Program &p) {
  std::map<std::string, uint32_t> m;
  for (uint32_t i = 0; i < p.ops.size(); i++) m[p.ops[i].name] = i;
  return m;
}

// executes one instruction on host copies of its constant inputs; `values` maps a byte address to
// the value stored there (components x lanes doubles)
inline void hostExec(const trxfile::Program &p, const DecodedInst &in,
                     std::unordered_map<uint64_t, std::vector<double>> &values) {
  uint16_t opcode;
  uint32_t lw;
  if (!trx::findOp(p.ops[in.op].name, opcode, lw)) throw std::runtime_error("operator not found: " + p.ops[in.op].name);
  const auto &op = p.ops[in.op];
  std::vector<double> scratch((size_t)in.n_args * 128, 0.0);
  uint32_t args[16];
  for (uint32_t k = 0; k < in.n_args; k++) {
    args[k] = k * 128;
    if (op.args[k].is_input) {
      auto &v = values.at(p.code[in.code_at + k]);
      std::memcpy(scratch.data() + args[k], v.data(), std::min<size_t>(v.size() * 8, op.args[k].size));
    }
  }
  for (int lane = 0; lane < (lw == 1 ? 1 : 8); lane++) {
    trxvm::Ctx c;
    c.m = scratch.data();
    c.ring = nullptr;
    c.nargs = (int)in.n_args;
    c.dual = false;
    c.rings = false;
    c.preloaded = false;
    c.gpool = nullptr;
    c.a = args;
    c.as = 1;
    c.cs = (lw == 1) ? 1 : 8;
    c.lo = lane;
    trxvm::exec(opcode, c);
  }
  for (uint32_t k = 0; k < in.n_args; k++)
    if (!op.args[k].is_input)
      values[p.code[in.code_at + k]].assign(scratch.begin() + args[k], scratch.begin() + args[k] + op.args[k].size / 8);
}

// recorder.cpp:117-199: an instruction whose inputs are all constants is evaluated now and its
// outputs become constants.
inline void foldConstants(trxfile::Program &p) {
  auto insts = decode(p);
  std::unordered_map<uint64_t, std::vector<double>> values;
  for (auto &c : p.constants) {
    std::vector<double> v(c.size / 8);
    std::memcpy(v.data(), p.const_data.data() + c.offset, c.size);
    values[c.address] = std::move(v);
  }
  std::vector<uint64_t> code;
  code.reserve(p.code.size());
  uint64_t kept = 0;
  auto meta = opMeta(p);
  for (auto &in : insts) {
    const auto &op = p.ops[in.op];
    bool is_const = !meta[in.op].sig->coop;
    for (uint32_t k = 0; k < in.n_args; k++)
      if (op.args[k].is_input && !values.count(p.code[in.code_at + k])) is_const = false;
    if (is_const) {
      hostExec(p, in, values);
      for (uint32_t k = 0; k < in.n_args; k++)
        if (!op.args[k].is_input) {
          uint64_t addr = p.code[in.code_at + k];
          trxfile::Port c;
          c.address = addr;
          c.offset = p.const_data.size();
          c.size = op.args[k].size;
          c.lanes = op.args[k].lanes;
          c.elem = op.args[k].elem;
          p.constants.push_back(c);
          const uint8_t *src = (const uint8_t *)values[addr].data();
          p.const_data.insert(p.const_data.end(), src, src + c.size);
        }
    } else {
      for (uint32_t k = 0; k < in.n_args; k++)
        if (!op.args[k].is_input) forEachKey(p, meta, in, k, [&](uint64_t a) { values.erase(a); });
      code.push_back(in.op);
      for (uint32_t k = 0; k < in.n_args; k++) code.push_back(p.code[in.code_at + k]);
      kept++;
    }
  }
  p.code.swap(code);
  p.n_instructions = kept;
}

// recorder.cpp:63-115 and :36-61
inline void removeUnused(trxfile::Program &p) {
  auto insts = decode(p);
  std::unordered_set<uint64_t> used;
  for (auto &o : p.outputs) used.insert(o.address);
  std::vector<uint8_t> keep(insts.size(), 0);
  auto meta = opMeta(p);
  for (size_t i = insts.size(); i-- > 0;) {
    auto &in = insts[i];
    const auto &op = p.ops[in.op];
    bool op_used = false;
    for (uint32_t k = 0; k < in.n_args; k++)
      if (!op.args[k].is_input) forEachKey(p, meta, in, k, [&](uint64_t a) { if (used.count(a)) op_used = true; });
    if (!op_used) continue;
    keep[i] = 1;
    for (uint32_t k = 0; k < in.n_args; k++)
      if (op.args[k].is_input) forEachKey(p, meta, in, k, [&](uint64_t a) { used.insert(a); });
  }
  std::vector<uint64_t> code;
  uint64_t n = 0;
  for (size_t i = 0; i < insts.size(); i++)
    if (keep[i]) {
      code.push_back(insts[i].op);
      for (uint32_t k = 0; k < insts[i].n_args; k++) code.push_back(p.code[insts[i].code_at + k]);
      n++;
    }
  p.code.swap(code);
  p.n_instructions = n;
  std::vector<trxfile::Port> consts;
  std::vector<uint8_t> data;
  for (auto &c : p.constants)
    if (used.count(c.address)) {
      trxfile::Port q = c;
      q.offset = data.size();
      data.insert(data.end(), p.const_data.begin() + c.offset, p.const_data.begin() + c.offset + c.size);
      consts.push_back(q);
    }
  p.constants.swap(consts);
  p.const_data.swap(data);
}

}  // namespace trxb
