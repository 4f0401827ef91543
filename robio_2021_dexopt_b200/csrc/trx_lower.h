// trx_lower.h -- lowering of a tractor tape to the B200 VM's instruction streams.
//
// What the reference does per program in SimpleEngine::_compile
// (tractor/src/engines/simple.cpp:7-20: flatten to {function pointer, argument
// base}) becomes here:
//   1. operator names -> device opcodes, argument layouts verified
//   2. dependency analysis over byte addresses (RAW / WAR / WAW) -> ASAP levels.
//      The only engine of the reference that reorders a tape is LoopEngine's
//      scheduleSimple (tractor/src/engines/loop.cpp:10-39); this is the same
//      levelisation, extended with WAR/WAW edges because bprop tapes overwrite
//      gradient slots (tractor/src/core/gradients.cpp:120-179).
//   3. same-opcode instructions of a level are packed into warp bundles:
//      4 lane-typed (*_8d) instructions x 8 lanes, or 32 scalar-typed (*_d)
//      instructions, per 32-thread warp, so a warp never diverges on the opcode
//   4. bundles are distributed over W per-warp streams (longest-processing-time
//      first), every level ends with a BARRIER word in every stream
// plus flat element maps for the port scatter/gather and constant upload
// kernels (ExecutableBase::_input/_output/_parameterize, src/engines/base.cpp:9-40;
// constant copy of SimpleEngine::_execute, simple.cpp:27-30).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/trx_file.h"
#include "trx_ops.h"

namespace trx {

// record header: opcode[0:11) | wide[11] | n_sub-1[12:17) | n_args[17:21) | coop[21]
//   bundle record:  header, then n_args x (4 | 32) argument words (slot index = byte address / 8)
//   coop record:    header, then n_args argument words and two immediates; present in every warp's stream
struct HeaderBits {
  static uint32_t make(uint32_t opcode, bool wide, uint32_t n_sub, uint32_t n_args, bool coop = false) {
    return (opcode & 0x7ffu) | ((wide ? 1u : 0u) << 11) | (((n_sub - 1) & 31u) << 12) | ((n_args & 15u) << 17) |
           ((coop ? 1u : 0u) << 21);
  }
};

// The wide layout places ports in order; a lane-typed port occupies comps * R elements.
// wide_base for element (port p, comp c, lane l) = fixed_before_p + per_lane_before_p * R + c * R + l,
// so it is stored as (fixed part, per-lane multiplier part) and resolved when R is known.
struct PortLayout {
  struct Elem {
    uint32_t idx;
    uint32_t fixed;     // scalar elements before this one
    uint32_t per_lane;  // lane-typed components before this one (multiplied by R)
    uint8_t lane;       // lane inside the tile (0..7) or 0xff for scalar-typed
  };
  std::vector<Elem> elems;
  uint64_t fixed_total = 0, per_lane_total = 0;
  uint64_t wideElems(uint64_t lanes) const { return fixed_total + per_lane_total * lanes; }
};

inline PortLayout makePortLayout(const std::vector<trxfile::Port> &ports, uint32_t lane_width, uint64_t memory_size) {
  PortLayout L;
  uint64_t fixed = 0, per_lane = 0, packed = 0;
  for (auto &p : ports) {
    if (p.address % 8 || p.size % 8) throw std::runtime_error("port not 8-byte aligned");
    // a foreign or damaged program must not make the scatter / gather kernels leave the tile image
    if (p.address + p.size > memory_size || p.address + p.size < p.address) throw std::runtime_error("port outside the memory image");
    // packed buffers are dense in declaration order (src/core/recorder.cpp:514-572, gradients.cpp:14-20)
    if (p.offset != packed) throw std::runtime_error("port offsets are not the packed offsets of the declaration order");
    packed += p.size;
    if (p.lanes == 1) {
      uint32_t n = p.size / 8;
      for (uint32_t c = 0; c < n; c++)
        L.elems.push_back({(uint32_t)(p.address / 8 + c), (uint32_t)(fixed + c), (uint32_t)per_lane, 0xff});
      fixed += n;
    } else {
      if (p.lanes != lane_width) throw std::runtime_error("port lane width differs from the tape's lane width");
      uint32_t comps = p.size / (8 * p.lanes);
      for (uint32_t c = 0; c < comps; c++)
        for (uint32_t l = 0; l < p.lanes; l++)
          L.elems.push_back({(uint32_t)(p.address / 8 + c * p.lanes + l), (uint32_t)fixed, (uint32_t)(per_lane + c),
                             (uint8_t)l});
      per_lane += comps;
    }
  }
  L.fixed_total = fixed;
  L.per_lane_total = per_lane;
  return L;
}

struct LowerOptions {
  uint32_t n_warps = 16;
  // software prefetch: operands of level L+D that already exist when level L starts (values of an earlier
  // program, constants, ports) are pulled into L2 while level L runs -- the schedule is static, so the
  // addresses are known.  0 = off.  Measured on the final kernel (profiles/README.md): -7 % per evaluation
  // at D = 1..4 (the adjoint reads the forward tape from DRAM); programs with a single level skip it.
  uint32_t prefetch_distance = 2;
  // CTA-wide on-chip slot cache (shared memory): a FIFO ring of the most recently produced lane-typed
  // values plus a pool of the tape's distinct constants.  Every result is still written to the memory
  // image (the reference's persist-everything contract); reads are served on chip when the lowering can
  // prove the value is still there.  0 = off.
  uint32_t ring_slots = 0;   // 64-byte slots
  uint32_t pool_slots = 0;   // 64-byte slots for deduplicated constants
  // dependent instructions whose only same-level producers sit in ONE lane-local instruction chain are
  // appended to that chain instead of opening a new level: the chain's records run back to back on the
  // same threads (thread = (sub-instruction, lane)), so program order replaces the block barrier.
  // Maximum chain length; <= 1 turns chaining off.
  uint32_t chain_cap = 4;  // measured on B200 (profiles/README.md): best for the folded adjoint; 5 and more lose to load imbalance
  // stream pages (words, power of two; 0 = unpaged): no record straddles a page boundary and every
  // warp's stream starts on one, so the kernel can stage whole pages in shared memory ahead of the
  // interpreter and read record headers and argument words from there
  uint32_t page_words = 512;
  // accumulate folding (adjoint tapes inside an objective, where nothing but ports crosses programs).
  // The reference accumulates a second gradient contribution as
  //     rop -> tmp ; move g -> t2 ; add t2, tmp -> g          (gradients.cpp:120-179)
  // because its operators never alias an output with an input.  Here every operator loads all inputs
  // before it stores, so the copy is dropped and the add runs in place:  add g, tmp -> g.  Likewise
  //     rop -> tmp ; move tmp -> g      becomes      rop -> g.
  // The dropped temporaries have one writer, one reader and are no port.  Values and summation order of
  // every surviving slot are unchanged.  (Folding the add itself into rop's store was measured slower:
  // the read-modify-write per component serialises behind rop's other stores.)
  bool fold_accumulate = false;
};

struct LoweredProgram {
  uint64_t memory_size = 0;
  uint32_t lane_width = 8;
  uint32_t n_warps = 8;
  std::vector<uint32_t> stream;
  std::vector<uint64_t> stream_begin;
  uint64_t n_instructions = 0, n_levels = 0, n_bundles = 0;
  uint64_t arg_bytes_per_tile = 0;
  uint64_t max_level_width = 0;
  uint32_t ring_doubles = 0;  // per-warp on-chip slot ring (doubles); 0 = every slot in the memory image
  uint64_t n_prefetch_lines = 0;
  uint32_t page_words = 0;  // stream page size the records were laid out for (0 = unpaged)
  uint64_t n_copies_forwarded = 0;  // reads redirected from a copy to its source (fold_accumulate)
  uint64_t n_folded = 0;   // move/add instructions folded into their producer's output (fold_accumulate)
  uint64_t n_dropped = 0;  // instructions without outputs, not lowered
  uint64_t n_chained = 0;  // instructions appended to a chain (each saved a level boundary)
  uint32_t global_pool = 0;             // 1: pool_values live in global memory (no on-chip ring)
  uint32_t cta_ring = 0;                // 1: ring_doubles is one CTA-wide cache (ring + constant pool), not per warp
  std::vector<double> pool_values;      // initial contents of the constant pool (doubles), placed after the ring
  uint32_t pool_offset_doubles = 0;
  std::vector<uint32_t> level_bundles;  // per level: bundles | (coop records << 16)  (development aid)
  uint64_t n_ring_reads = 0, n_image_reads = 0, n_ring_only_writes = 0, n_dual_writes = 0, n_image_writes = 0;
  bool has_forward = false, has_reverse = false;
  bool draws_random = false;  // add_random_* / dropout2 on the tape: the executor sets the generator state per execute
  bool uses_shapes = false;   // collision_axes / collision_project_2 on the tape: the executor mirrors the shape table
  PortLayout inputs, parameters, outputs;
  std::vector<uint32_t> const_idx;
  std::vector<uint64_t> const_bits;
};

inline void lowerProgram(const trxfile::Program &src, const LowerOptions &opt, LoweredProgram &out) {
  const auto &table = opTable();
  out = LoweredProgram();
  out.memory_size = (src.memory_size + 255) / 256 * 256;
  out.n_warps = opt.n_warps;
  if (opt.n_warps < 1 || opt.n_warps > 20) throw std::runtime_error("n_warps out of range (1..20)");
  if (src.memory_size / 8 >= 0xffffffffull) throw std::runtime_error("memory image too large for 32-bit slot indices");

  // ---- 1. operator table -> opcodes, signature check
  struct OpRef {
    uint16_t opcode;
    uint32_t lane_width;
  };
  std::vector<OpRef> ops(src.ops.size());
  uint32_t tape_lanes = 1;
  for (size_t i = 0; i < src.ops.size(); i++) {
    const auto &d = src.ops[i];
    uint16_t opcode;
    uint32_t lw;
    if (!findOp(d.name, opcode, lw))
      throw std::runtime_error("unsupported operator on the device engine (no CPU fallback): " + d.name);
    const OpSig &sig = table[opcode];
    if (sig.args.size() != d.args.size())
      throw std::runtime_error("argument count mismatch for " + d.name);
    for (size_t k = 0; k < d.args.size(); k++) {
      const ArgSig &a = sig.args[k];
      if (isBlock(a.kind)) {
        uint32_t eb = blockElemBytes(a.kind, 8);
        if (d.args[k].size == 0 || d.args[k].size % eb || (d.args[k].is_input != 0) != a.is_input)
          throw std::runtime_error("block argument layout mismatch for " + d.name + " arg " + std::to_string(k));
        continue;
      }
      uint32_t expect_lanes = (a.kind == ARG_T) ? lw : 1;
      uint32_t expect_size = a.comps * 8 * expect_lanes;
      if (d.args[k].size != expect_size || (d.args[k].is_input != 0) != a.is_input || d.args[k].lanes != expect_lanes)
        throw std::runtime_error("argument layout mismatch for " + d.name + " arg " + std::to_string(k));
    }
    if (sig.coop) tape_lanes = std::max<uint32_t>(tape_lanes, 8);
    ops[i] = {opcode, lw};
    tape_lanes = std::max(tape_lanes, lw);
    const std::string base_name = sig.base;  // (a string, not the table's pointer)
    if (sig.mode == 'C' && (base_name == "add_random_normal" || base_name == "add_random_uniform" || base_name == "dropout2"))
      out.draws_random = true;
    if (sig.mode == 'C' && (base_name == "collision_axes" || base_name == "collision_project_2")) out.uses_shapes = true;
    if (sig.mode == 'F') out.has_forward = true;
    if (sig.mode == 'R') out.has_reverse = true;
  }
  for (auto *ports : {&src.inputs, &src.parameters, &src.outputs, &src.constants})
    for (auto &p : *ports) tape_lanes = std::max<uint32_t>(tape_lanes, p.lanes);
  if (tape_lanes != 1 && tape_lanes != 8) throw std::runtime_error("only Batch<double,8> lane tapes are supported");
  out.lane_width = 8;  // scalar-only tapes still run on 8-lane tiles (lane 0 active)

  // ---- 2. decode instructions
  struct Inst {
    uint32_t op;      // index into ops
    uint64_t code_at; // position of the first argument in src.code
    uint32_t level;
    uint32_t chain;   // chain id (kNoChain: not chainable), position = order of appearance in the level
  };
  constexpr uint32_t kNoChain = 0xffffffffu;
  std::vector<Inst> insts;
  insts.reserve(src.n_instructions);
  uint64_t n_decoded = 0;
  {
    uint64_t pc = 0;
    while (pc < src.code.size()) {
      uint64_t oi = src.code[pc];
      if (oi >= ops.size()) throw std::runtime_error("invalid op");  // engine.cpp:44-48
      size_t na = table[ops[oi].opcode].args.size();
      if (pc + 1 + na > src.code.size()) throw std::runtime_error("truncated instruction");
      // an instruction without an output argument has no effect (the prepare rules of operators whose
      // reverse rule needs nothing saved, e.g. prepare_add): it is counted but not lowered
      bool has_output = false;
      for (auto &a : table[ops[oi].opcode].args) has_output |= !a.is_input;
      n_decoded++;
      if (has_output) insts.push_back({(uint32_t)oi, pc + 1, 0, kNoChain});
      else out.n_dropped++;
      pc += 1 + na;
    }
    if (n_decoded != src.n_instructions) throw std::runtime_error("instruction count mismatch");
  }
  out.n_instructions = n_decoded;

  // The dense-layer kernels read their weight and bias blocks through the read-only data path (__ldg): that is only
  // valid while no EARLIER instruction of the same launch has written those slots (a later writer -- the adjoint of
  // an hprop program accumulating into the slots its tangent sweep read -- is ordered behind the read by a level
  // barrier).  Weights are input ports in every tape the recorders produce; a program that computes them first is
  // refused rather than served stale values.
  {
    std::vector<std::pair<uint64_t, uint64_t>> written;  // [begin, end) of the output arguments so far (program order)
    for (auto &in : insts) {
      const OpSig &sig = table[ops[in.op].opcode];
      if (sig.coop)
        for (size_t k = 0; k < sig.args.size(); k++) {
          if (!sig.args[k].is_input || sig.args[k].kind != ARG_SB) continue;
          const uint64_t a = src.code[in.code_at + k], e = a + src.ops[in.op].args[k].size;
          for (auto &w : written)
            if (w.first < e && w.second > a)
              throw std::runtime_error("unsupported operator use: a dense layer's weight block is computed by the same program");
        }
      for (size_t k = 0; k < sig.args.size(); k++)
        if (!sig.args[k].is_input) {
          const uint64_t a = src.code[in.code_at + k];
          // (only writers of scalar slots can alias a weight block; lane-typed outputs live elsewhere)
          if (src.ops[in.op].args[k].lanes == 1 || sig.coop) written.push_back({a, a + src.ops[in.op].args[k].size});
        }
    }
  }

  // ---- 2b. copy propagation and accumulate folding (see LowerOptions)
  std::vector<uint64_t> code(src.code);        // argument addresses after the rewrites
  std::vector<uint64_t> base;                  // ... after copy propagation: what the folding patterns match on
  if (opt.fold_accumulate && opt.ring_slots == 0 && opt.pool_slots == 0) {
    // -- copy propagation.  An output that is a plain copy of an input (move; the gradient copies of
    // reverse_move and reverse_add) is forwarded to its readers: they read the source instead, as long as
    // neither slot was written in between, and a copy instruction whose every output lost all its readers is
    // dropped.  This also turns the reference's  `move g -> t2 ; add t2, tmp -> g`  into the in-place add.
    {
      struct CopyOf {
        uint64_t source;
        uint32_t source_version, own_version, inst, k_out, bytes;
      };
      struct CopyRule {
        int k_out, k_in;
      };
      auto copyRules = [&](uint16_t o, CopyRule *r) -> int {
        switch (o) {
          case OP_MOVE: case OP_MOVE_MAT3: case OP_MOVE_VEC3: case OP_MOVE_QUAT: case OP_MOVE_POSE: case OP_MOVE_TWIST:
            r[0] = {1, 0};
            return 1;
          case OP_R_MOVE: case OP_R_MOVE_MAT3: case OP_R_MOVE_VEC3: case OP_R_MOVE_QUAT: case OP_R_MOVE_POSE: case OP_R_MOVE_TWIST:
            r[0] = {0, 1};
            return 1;
          case OP_R_ADD: case OP_R_ADD_VEC3: case OP_R_ADD_TWIST:
            r[0] = {0, 2};
            r[1] = {1, 2};
            return 2;
          default:
            return 0;
        }
      };
      std::unordered_map<uint64_t, uint32_t> version, n_writes, n_reads;
      version.reserve(src.code.size());
      std::unordered_set<uint64_t> port_slots;
      for (auto *ports : {&src.inputs, &src.parameters, &src.outputs, &src.constants})
        for (auto &p : *ports)
          for (uint64_t a = p.address; a < p.address + p.size; a += 8) port_slots.insert(a);
      auto forEachSlot = [&](const Inst &in, size_t k, const std::function<void(uint64_t)> &f) {
        const ArgSig &a = table[ops[in.op].opcode].args[k];
        const uint64_t addr = src.code[in.code_at + k];
        if (!isBlock(a.kind)) {
          f(addr);
          return;
        }
        const uint32_t eb = blockElemBytes(a.kind, 8), size = src.ops[in.op].args[k].size;
        for (uint32_t off = 0; off < size; off += eb) f(addr + off);
      };
      for (auto &in : insts) {
        const OpSig &sig = table[ops[in.op].opcode];
        for (size_t k = 0; k < sig.args.size(); k++)
          forEachSlot(in, k, [&](uint64_t a) { (sig.args[k].is_input ? n_reads : n_writes)[a]++; });
      }
      std::unordered_map<uint64_t, CopyOf> copy_of;
      std::vector<std::array<uint32_t, 2>> redirected(insts.size(), std::array<uint32_t, 2>{0, 0});
      std::vector<uint8_t> is_copy(insts.size(), 0);
      for (uint32_t i = 0; i < insts.size(); i++) {
        const Inst &in = insts[i];
        const OpRef &o = ops[in.op];
        const OpSig &sig = table[o.opcode];
        // readers first: an instruction reads its inputs before it writes
        if (!sig.coop)
          for (size_t k = 0; k < sig.args.size(); k++) {
            if (!sig.args[k].is_input || sig.args[k].kind != ARG_T) continue;
            auto c = copy_of.find(code[in.code_at + k]);
            if (c == copy_of.end()) continue;
            const CopyOf &co = c->second;
            if (version[co.source] != co.source_version || version[c->first] != co.own_version) continue;
            if (co.bytes != src.ops[in.op].args[k].size) continue;
            code[in.code_at + k] = co.source;
            redirected[co.inst][co.k_out]++;
            out.n_copies_forwarded++;
          }
        for (size_t k = 0; k < sig.args.size(); k++)
          if (!sig.args[k].is_input) forEachSlot(in, k, [&](uint64_t a) { version[a]++; });
        CopyRule rules[2];
        const int nr = copyRules(o.opcode, rules);
        for (int q = 0; q < nr; q++) {
          const uint64_t t = src.code[in.code_at + rules[q].k_out], s_ = code[in.code_at + rules[q].k_in];
          if (t == s_) continue;
          is_copy[i] = 1;
          copy_of[t] = {s_, version[s_], version[t], i, (uint32_t)q, src.ops[in.op].args[rules[q].k_out].size};
        }
      }
      // drop the copy instructions nobody reads any more
      size_t w = 0;
      for (uint32_t i = 0; i < insts.size(); i++) {
        bool drop = false;
        if (is_copy[i]) {
          const Inst &in = insts[i];
          CopyRule rules[2];
          const int nr = copyRules(ops[in.op].opcode, rules);
          drop = true;
          for (int q = 0; q < nr && drop; q++) {
            const uint64_t t = src.code[in.code_at + rules[q].k_out];
            drop = n_writes[t] == 1 && !port_slots.count(t) && redirected[i][q] == n_reads[t] &&
                   t != code[in.code_at + rules[q].k_in];
          }
          // both outputs of reverse_add into the same slot would make the counts ambiguous: keep it
          if (nr == 2 && src.code[in.code_at + rules[0].k_out] == src.code[in.code_at + rules[1].k_out]) drop = false;
        }
        if (drop) out.n_folded++;
        else insts[w++] = insts[i];
      }
      insts.resize(w);
    }
    base = code;
    struct SlotUse {
      uint32_t writer = 0, reader = 0;
      uint32_t nw = 0, nr = 0;
    };
    std::unordered_map<uint64_t, SlotUse> use;
    use.reserve(base.size());
    for (uint32_t i = 0; i < insts.size(); i++) {
      const Inst &in = insts[i];
      const OpSig &sig = table[ops[in.op].opcode];
      for (size_t k = 0; k < sig.args.size(); k++) {
        uint64_t addr = base[in.code_at + k];
        if (isBlock(sig.args[k].kind)) {
          uint32_t eb = blockElemBytes(sig.args[k].kind, 8), size = src.ops[in.op].args[k].size;
          for (uint32_t off = 0; off < size; off += eb) {
            SlotUse &u = use[addr + off];
            u.nw += 2;
            u.nr += 2;
          }
          continue;
        }
        SlotUse &u = use[addr];
        if (sig.args[k].is_input) {
          u.nr++;
          u.reader = i;
        } else {
          u.nw++;
          u.writer = i;
        }
      }
    }
    std::vector<std::pair<uint64_t, uint64_t>> port_ranges;
    for (auto *ports : {&src.inputs, &src.parameters, &src.outputs, &src.constants})
      for (auto &p : *ports) port_ranges.push_back({p.address, p.address + p.size});
    std::sort(port_ranges.begin(), port_ranges.end());
    {  // merge overlapping ranges: afterwards only the range starting last before the query can overlap it
      size_t w = 0;
      for (size_t i = 0; i < port_ranges.size(); i++) {
        if (w && port_ranges[i].first <= port_ranges[w - 1].second)
          port_ranges[w - 1].second = std::max(port_ranges[w - 1].second, port_ranges[i].second);
        else
          port_ranges[w++] = port_ranges[i];
      }
      port_ranges.resize(w);
    }
    auto inPort = [&](uint64_t a, uint64_t size) {
      auto it = std::lower_bound(port_ranges.begin(), port_ranges.end(), std::make_pair(a + size, (uint64_t)0));
      if (it == port_ranges.begin()) return false;
      --it;
      return it->second > a;
    };
    auto isMove = [](uint16_t o) {
      return o == OP_MOVE || o == OP_MOVE_MAT3 || o == OP_MOVE_VEC3 || o == OP_MOVE_QUAT || o == OP_MOVE_POSE || o == OP_MOVE_TWIST;
    };
    auto isAdd = [](uint16_t o) { return o == OP_ADD || o == OP_ADD_VEC3 || o == OP_ADD_TWIST; };
    std::vector<uint8_t> dropped(insts.size(), 0);
    // a slot with exactly one writer and one reader (instruction `reader`): returns the writer or ~0u
    auto soleWriter = [&](uint64_t addr, uint32_t reader) -> uint32_t {
      auto it = use.find(addr);
      if (it == use.end() || it->second.nw != 1 || it->second.nr != 1 || it->second.reader != reader) return ~0u;
      return it->second.writer;
    };
    for (uint32_t bi = 0; bi < insts.size(); bi++) {
      const Inst &B = insts[bi];
      const OpRef &bo = ops[B.op];
      if (!isMove(bo.opcode) && !isAdd(bo.opcode)) continue;
      const uint32_t comps = table[bo.opcode].args[0].comps;
      const uint64_t bytes = (uint64_t)comps * 8 * bo.lane_width;
      // nothing in (from, bi) other than `skip` may touch dst
      auto untouched = [&](uint32_t from, uint32_t skip, uint64_t dst) {
        for (uint32_t j = from + 1; j < bi; j++) {
          if (j == skip) continue;
          const Inst &J = insts[j];
          const OpSig &jsig = table[ops[J.op].opcode];
          if (jsig.coop) return false;
          for (size_t k = 0; k < jsig.args.size(); k++)
            if (code[J.code_at + k] == dst || base[J.code_at + k] == dst) return false;
        }
        return true;
      };
      if (isAdd(bo.opcode)) {
        // move dst -> t2 ; add t2, tmp -> dst   becomes the in-place   add dst, tmp -> dst
        // (every operator loads all its inputs before it stores; vm_ops.def)
        const uint64_t dst = base[B.code_at + 2];
        for (int q = 0; q < 2; q++) {
          const uint64_t t2 = base[B.code_at + q], other = base[B.code_at + 1 - q];
          if (t2 == dst || other == dst || other == t2) continue;
          const uint32_t mi = soleWriter(t2, bi);
          if (mi == ~0u || mi >= bi || bi - mi > 16 || dropped[mi] || inPort(t2, bytes)) continue;
          const Inst &M = insts[mi];
          const OpRef &mo = ops[M.op];
          if (!isMove(mo.opcode) || mo.lane_width != bo.lane_width || table[mo.opcode].args[0].comps != comps) continue;
          if (base[M.code_at] != dst || base[M.code_at + 1] != t2) continue;
          if (!untouched(mi, ~0u, dst)) continue;
          code[B.code_at + q] = dst;
          dropped[mi] = 1;
          out.n_folded++;
          break;
        }
        continue;
      }
      // A -> tmp ; move tmp -> dst   becomes   A -> dst
      const uint64_t tmp = base[B.code_at], dst = base[B.code_at + 1];
      if (tmp == dst) continue;
      const uint32_t ai = soleWriter(tmp, bi);
      if (ai == ~0u || ai >= bi || bi - ai > 16 || dropped[ai]) continue;
      const Inst &A = insts[ai];
      const OpRef &ao = ops[A.op];
      const OpSig &asig = table[ao.opcode];
      if (asig.coop || ao.lane_width != bo.lane_width) continue;
      bool local = true;
      for (auto &a : asig.args) local &= (a.kind == ARG_T) || (a.kind == ARG_S && ao.lane_width == 1);
      if (!local) continue;
      int k_out = -1;
      bool dst_used = false;
      for (size_t k = 0; k < asig.args.size(); k++) {
        if (code[A.code_at + k] == dst) dst_used = true;  // A reads dst or already writes it
        if (asig.args[k].is_input) continue;
        if (base[A.code_at + k] == tmp && asig.args[k].comps == comps && asig.args[k].kind == ARG_T) k_out = (int)k;
      }
      if (k_out < 0 || dst_used || inPort(tmp, bytes) || !untouched(ai, ~0u, dst)) continue;
      code[A.code_at + k_out] = dst;
      dropped[bi] = 1;
      out.n_folded++;
    }
    if (out.n_folded) {
      size_t w = 0;
      for (size_t i = 0; i < insts.size(); i++)
        if (!dropped[i]) insts[w++] = insts[i];
      insts.resize(w);
    }
  }

  // ---- 3. levels (ASAP) with RAW / WAR / WAW edges keyed by argument start address
  // an argument is tracked by its start address; block arguments by the start address of every element
  auto forEachKey = [&](const Inst &in, size_t k, const std::function<void(uint64_t)> &f) {
    const ArgSig &a = table[ops[in.op].opcode].args[k];
    uint64_t addr = code[in.code_at + k];
    if (!isBlock(a.kind)) {
      f(addr);
      return;
    }
    uint32_t eb = blockElemBytes(a.kind, 8), size = src.ops[in.op].args[k].size;
    for (uint32_t off = 0; off < size; off += eb) f(addr + off);
  };
  std::vector<uint64_t> addrs;
  addrs.reserve(src.code.size());
  for (auto &in : insts) {
    size_t na = table[ops[in.op].opcode].args.size();
    for (size_t k = 0; k < na; k++) {
      uint64_t a = code[in.code_at + k];
      if (a % 8) throw std::runtime_error("argument address not 8-byte aligned");
      if (a + src.ops[in.op].args[k].size > src.memory_size)
        throw std::runtime_error("argument address outside the memory image");
      forEachKey(in, k, [&](uint64_t key) { addrs.push_back(key); });
    }
  }
  std::sort(addrs.begin(), addrs.end());
  addrs.erase(std::unique(addrs.begin(), addrs.end()), addrs.end());
  auto slotId = [&](uint64_t a) -> size_t {
    return (size_t)(std::lower_bound(addrs.begin(), addrs.end(), a) - addrs.begin());
  };
  std::vector<uint32_t> wlevel(addrs.size(), 0), rlevel(addrs.size(), 0);
  // chain of the last writer / of the readers at rlevel (kNoChain when they belong to different chains)
  const bool chaining = opt.chain_cap > 1 && opt.ring_slots == 0;
  std::vector<uint32_t> wchain(chaining ? addrs.size() : 0, kNoChain), rchain(chaining ? addrs.size() : 0, kNoChain);
  std::vector<uint32_t> chain_len;   // per chain
  std::vector<uint8_t> chain_wide;   // per chain: scalar-typed (32 per warp) or lane-typed (4 x 8 lanes)
  // lane-local: thread (sub-instruction, lane) touches only its own lane of every argument
  auto laneLocal = [&](const OpRef &o, const OpSig &sig) {
    if (sig.coop) return false;
    for (auto &a : sig.args) {
      if (a.kind == ARG_T) continue;
      if (a.kind == ARG_S && o.lane_width == 1) continue;
      return false;
    }
    return true;
  };
  uint32_t n_levels = 0;
  // level of the last writer of each (non-block) input argument at the time it is read: decides
  // whether the operand already exists when a prefetch for it could be issued
  std::vector<uint32_t> arg_wlevel(src.code.size(), 0);
  for (auto &in : insts) {
    const OpSig &sig = table[ops[in.op].opcode];
    size_t na = sig.args.size();
    uint32_t lvl = 0;
    uint32_t dep_chain = kNoChain;
    bool dep_mixed = false;  // the dependencies at level `lvl` are not all in one chain
    auto dep = [&](uint32_t l, uint32_t c) {
      if (l == 0) return;
      if (l > lvl) {
        lvl = l;
        dep_chain = c;
        dep_mixed = (c == kNoChain);
      } else if (l == lvl && (c != dep_chain || c == kNoChain)) {
        dep_mixed = true;
      }
    };
    for (size_t k = 0; k < na; k++) {
      bool is_in = sig.args[k].is_input;
      uint32_t wl = 0;
      forEachKey(in, k, [&](uint64_t key) {
        size_t id = slotId(key);
        wl = std::max(wl, wlevel[id]);
        dep(wlevel[id], chaining ? wchain[id] : kNoChain);
        if (!is_in) dep(rlevel[id], chaining ? rchain[id] : kNoChain);
      });
      arg_wlevel[in.code_at + k] = wl;
      out.arg_bytes_per_tile += src.ops[in.op].args[k].size;
    }
    const bool local = chaining && laneLocal(ops[in.op], sig);
    const bool wide_op = ops[in.op].lane_width == 1;
    if (local && lvl >= 1 && !dep_mixed && dep_chain != kNoChain && chain_wide[dep_chain] == (wide_op ? 1 : 0) &&
        chain_len[dep_chain] < opt.chain_cap) {
      in.level = lvl;
      in.chain = dep_chain;
      chain_len[dep_chain]++;
      out.n_chained++;
    } else {
      in.level = lvl + 1;
      if (local) {
        in.chain = (uint32_t)chain_len.size();
        chain_len.push_back(1);
        chain_wide.push_back(wide_op ? 1 : 0);
      }
    }
    n_levels = std::max(n_levels, in.level);
    for (size_t k = 0; k < na; k++)
      if (sig.args[k].is_input)
        forEachKey(in, k, [&](uint64_t key) {
          size_t id = slotId(key);
          if (in.level > rlevel[id]) {
            rlevel[id] = in.level;
            if (chaining) rchain[id] = in.chain;
          } else if (chaining && in.level == rlevel[id] && rchain[id] != in.chain) {
            rchain[id] = kNoChain;
          }
        });
    for (size_t k = 0; k < na; k++)
      if (!sig.args[k].is_input)
        forEachKey(in, k, [&](uint64_t key) {
          size_t id = slotId(key);
          wlevel[id] = in.level;
          rlevel[id] = in.level;
          if (chaining) wchain[id] = rchain[id] = in.chain;
        });
  }
  out.n_levels = n_levels;
  out.page_words = opt.page_words;

  // ---- 4. bucket by level, then by (opcode, width), build bundles, distribute to warps
  std::vector<uint32_t> level_count(n_levels + 2, 0);
  for (auto &in : insts) level_count[in.level + 1]++;
  for (size_t l = 1; l < level_count.size(); l++) level_count[l] += level_count[l - 1];
  std::vector<uint32_t> order(insts.size());
  {
    std::vector<uint32_t> fill(level_count.begin(), level_count.end());
    for (uint32_t i = 0; i < insts.size(); i++) order[fill[insts[i].level]++] = i;
  }
  const uint32_t W = opt.n_warps;
  std::vector<std::vector<uint32_t>> streams(W);
  const uint32_t PG = opt.page_words;
  if (PG && ((PG & (PG - 1)) || PG < 512)) throw std::runtime_error("stream page size must be a power of two >= 512");
  // make room for an n-word record: it must not straddle a page
  auto room = [&](std::vector<uint32_t> &s, size_t n) {
    if (!PG) return;
    if (n > PG) throw std::runtime_error("record larger than a stream page");
    if ((s.size() & (PG - 1)) + n > PG)
      while (s.size() & (PG - 1)) s.push_back(HeaderBits::make(OP_PAGE, false, 1, 0));
  };
  // 128-byte lines (index = byte address / 128) that level L reads and that exist `prefetch_distance`
  // levels earlier
  const uint32_t PD = n_levels > 1 ? opt.prefetch_distance : 0;
  std::vector<std::vector<uint32_t>> level_lines(PD ? n_levels + 2 : 0);
  if (PD)
    for (auto &in : insts) {
      const OpSig &sig = table[ops[in.op].opcode];
      if (sig.coop) continue;
      // the prefetch for level L is issued when level P = max(1, L - PD) starts; the operand must be
      // complete by then (writer level < P; 0 = not written by this program: constants, ports, state of
      // an earlier program)
      const uint32_t P = in.level > PD ? in.level - PD : 1;
      for (size_t k = 0; k < sig.args.size(); k++) {
        if (!sig.args[k].is_input) continue;
        if (arg_wlevel[in.code_at + k] >= P) continue;
        uint64_t a = code[in.code_at + k], e = a + src.ops[in.op].args[k].size - 1;
        for (uint64_t line = a / 128; line <= e / 128; line++) level_lines[in.level].push_back((uint32_t)line);
      }
    }
  auto emitPrefetch = [&](uint32_t for_level) {
    if (!PD || for_level > n_levels) return;
    auto &lines = level_lines[for_level];
    if (lines.empty()) return;
    std::sort(lines.begin(), lines.end());
    lines.erase(std::unique(lines.begin(), lines.end()), lines.end());
    uint32_t w = 0;
    for (size_t i = 0; i < lines.size(); i += 32) {
      uint32_t n = (uint32_t)std::min<size_t>(32, lines.size() - i);
      auto &s = streams[w];
      room(s, 33);
      s.push_back(HeaderBits::make(OP_PREFETCH, true, n, 0));
      for (uint32_t j = 0; j < 32; j++) s.push_back(lines[i + (j < n ? j : 0)]);
      w = (w + 1) % W;
      out.n_prefetch_lines += n;
    }
    std::vector<uint32_t>().swap(lines);
  };
  for (uint32_t l = 1; l <= PD && l <= n_levels; l++) emitPrefetch(l);  // nothing runs before level 1: warm up
  struct Bundle {
    bool wide;
    uint32_t first, count;  // range in the level's sorted unit list
    uint32_t len;           // records: one per chain position
    uint32_t cost;
  };
  std::vector<uint32_t> unit_first, unit_order;
  std::vector<uint32_t> lvl_insts;
  std::vector<Bundle> bundles;
  std::vector<uint64_t> load(W);

  // ---- on-chip slot cache (see LowerOptions): static residency decisions, level by level
  const bool use_cache = opt.ring_slots > 0;
  const bool use_gpool = !use_cache && opt.pool_slots > 0;  // constant pool in global memory, no ring
  struct RingState {
    std::vector<int64_t> owner;                                          // per slot: byte address or -1
    std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> where;   // address -> (slot, comps)
    uint32_t head = 0;
    void drop(uint64_t addr) {
      auto it = where.find(addr);
      if (it == where.end()) return;
      for (uint32_t t = it->second.first; t < it->second.first + it->second.second; t++) owner[t] = -1;
      where.erase(it);
    }
    uint32_t alloc(uint64_t addr, uint32_t comps) {
      uint32_t n = (uint32_t)owner.size();
      if (comps > n) return ~0u;
      drop(addr);
      if (head + comps > n) head = 0;
      for (uint32_t t = head; t < head + comps; t++)
        if (owner[t] >= 0) drop((uint64_t)owner[t]);
      for (uint32_t t = head; t < head + comps; t++) owner[t] = (int64_t)addr;
      where[addr] = {head, comps};
      uint32_t r = head;
      head += comps;
      return r;
    }
  } ring;
  ring.owner.assign(opt.ring_slots, -1);
  std::unordered_map<uint64_t, uint32_t> pool_of_const;  // constant address -> pool slot
  if (use_cache || use_gpool) {
    out.cta_ring = use_cache ? 1 : 0;
    out.global_pool = use_gpool ? 1 : 0;
    out.pool_offset_doubles = use_cache ? opt.ring_slots * 8 : 0;
    // distinct lane-typed constants of this program, first come first served
    std::map<std::vector<uint64_t>, uint32_t> seen;
    std::unordered_set<uint64_t> written;
    for (auto &in : insts) {
      const OpSig &sig = table[ops[in.op].opcode];
      for (size_t k = 0; k < sig.args.size(); k++)
        if (!sig.args[k].is_input) forEachKey(in, k, [&](uint64_t key) { written.insert(key); });
    }
    uint32_t used = 0;
    for (auto &c : src.constants) {
      if (c.lanes != 8 || c.size % 64 || written.count(c.address)) continue;
      uint32_t comps = c.size / 64;
      std::vector<uint64_t> bits(c.size / 8);
      std::memcpy(bits.data(), src.const_data.data() + c.offset, c.size);
      auto it = seen.find(bits);
      if (it == seen.end()) {
        if (used + comps > opt.pool_slots) continue;
        it = seen.emplace(bits, used).first;
        for (uint64_t b : bits) {
          double d;
          std::memcpy(&d, &b, 8);
          out.pool_values.push_back(d);
        }
        used += comps;
      }
      pool_of_const[c.address] = it->second;
    }
    if (use_cache) out.ring_doubles = opt.ring_slots * 8 + (uint32_t)out.pool_values.size();
  }
  auto ringEligible = [&](const OpRef &o, const ArgSig &a) { return o.lane_width == 8 && a.kind == ARG_T && a.comps > 0; };

  for (uint32_t level = 1; level <= n_levels; level++) {
    emitPrefetch(level + PD);
    lvl_insts.assign(order.begin() + level_count[level], order.begin() + level_count[level + 1]);
    out.max_level_width = std::max<uint64_t>(out.max_level_width, lvl_insts.size());
    std::stable_sort(lvl_insts.begin(), lvl_insts.end(), [&](uint32_t a, uint32_t b) {
      const OpRef &oa = ops[insts[a].op], &ob = ops[insts[b].op];
      if (oa.opcode != ob.opcode) return oa.opcode < ob.opcode;
      return oa.lane_width < ob.lane_width;
    });
    bundles.clear();
    uint32_t n_coop_level = 0;
    // block-cooperative instructions first: one record, replicated into every warp's stream
    for (size_t i = 0; i < lvl_insts.size(); i++) {
      const Inst &in = insts[lvl_insts[i]];
      const OpRef &o = ops[in.op];
      const OpSig &sig = table[o.opcode];
      if (!sig.coop) continue;
      const auto &d = src.ops[in.op];
      uint32_t imm0 = 0, imm1 = 0;
      if (o.opcode == OP_X_DENSE || o.opcode == OP_R_X_DENSE || o.opcode == OP_F_X_DENSE) {
        imm0 = d.args[0].size / 64;
        imm1 = d.args[3].size / 64;
        if (d.args[1].size != (uint64_t)imm0 * imm1 * 8 || d.args[2].size != imm1 * 8)
          throw std::runtime_error("dense layer block sizes are inconsistent");
      } else if (o.opcode == OP_X_ZERO_BLOCK) {
        imm0 = d.args[0].size / 8;
      } else {
        throw std::runtime_error("unknown block operator " + d.name);
      }
      uint32_t na = (uint32_t)sig.args.size();
      if (use_cache)
        for (uint32_t k = 0; k < na; k++)
          if (!sig.args[k].is_input) forEachKey(in, k, [&](uint64_t key) { ring.drop(key); });
      for (auto &s : streams) {
        room(s, na + 3);
        s.push_back(HeaderBits::make(o.opcode, false, 1, na, true));
        for (uint32_t k = 0; k < na; k++) s.push_back((uint32_t)(code[in.code_at + k] / 8));
        s.push_back(imm0);
        s.push_back(imm1);
      }
      out.n_bundles++;
      n_coop_level++;
    }
    lvl_insts.erase(std::remove_if(lvl_insts.begin(), lvl_insts.end(),
                                   [&](uint32_t i) { return table[ops[insts[i].op].opcode].coop; }),
                    lvl_insts.end());
    // units = chains (or single instructions), members in program order; units with the same opcode
    // sequence are packed side by side: 4 lane-typed or 32 scalar-typed units per warp
    {
      std::vector<std::pair<uint64_t, uint32_t>> keyed(lvl_insts.size());
      for (size_t i = 0; i < lvl_insts.size(); i++) {
        uint32_t c = insts[lvl_insts[i]].chain;
        keyed[i] = {c == kNoChain ? (1ull << 32) + i : (uint64_t)c, lvl_insts[i]};
      }
      std::sort(keyed.begin(), keyed.end());  // by chain, then program order (instruction index)
      unit_first.clear();
      for (size_t i = 0; i < keyed.size(); i++) {
        if (i == 0 || keyed[i].first != keyed[i - 1].first) unit_first.push_back((uint32_t)i);
        lvl_insts[i] = keyed[i].second;
      }
      unit_first.push_back((uint32_t)keyed.size());
    }
    const size_t n_units = unit_first.size() - 1;
    auto unitLen = [&](uint32_t u) { return unit_first[u + 1] - unit_first[u]; };
    auto sigKey = [&](uint32_t u, uint32_t p) {
      const OpRef &o = ops[insts[lvl_insts[unit_first[u] + p]].op];
      return ((uint32_t)o.opcode << 4) | o.lane_width;
    };
    unit_order.resize(n_units);
    for (uint32_t u = 0; u < n_units; u++) unit_order[u] = u;
    auto unitCmp = [&](uint32_t x, uint32_t y) -> int {
      uint32_t lx = unitLen(x), ly = unitLen(y);
      for (uint32_t p = 0; p < lx && p < ly; p++) {
        uint32_t kx = sigKey(x, p), ky = sigKey(y, p);
        if (kx != ky) return kx < ky ? -1 : 1;
      }
      return lx == ly ? 0 : (lx < ly ? -1 : 1);
    };
    std::stable_sort(unit_order.begin(), unit_order.end(), [&](uint32_t x, uint32_t y) { return unitCmp(x, y) < 0; });
    for (size_t i = 0; i < n_units;) {
      uint32_t u = unit_order[i];
      const OpRef &o = ops[insts[lvl_insts[unit_first[u]]].op];
      bool wide = (o.lane_width == 1);
      uint32_t cap = wide ? 32 : 4;
      size_t j = i + 1;
      while (j < n_units && j - i < cap && unitCmp(u, unit_order[j]) == 0) j++;
      uint32_t cost = 0;
      for (uint32_t p = 0; p < unitLen(u); p++) cost += table[ops[insts[lvl_insts[unit_first[u] + p]].op].opcode].cost + 8;
      bundles.push_back({wide, (uint32_t)i, (uint32_t)(j - i), unitLen(u), cost});
      out.n_bundles += unitLen(u);
      i = j;
    }
    // instruction of the bundle's i-th unit at chain position p
    auto member = [&](const Bundle &bd, uint32_t i, uint32_t p) -> const Inst & {
      return insts[lvl_insts[unit_first[unit_order[bd.first + i]] + p]];
    };
    // results of this level go into the ring first (evicting the oldest values); the operands of this
    // level are resolved afterwards, so a value evicted now is read from the memory image
    std::unordered_map<uint64_t, uint32_t> out_slot;
    if (use_cache)
      for (uint32_t ii : lvl_insts) {
        const Inst &in = insts[ii];
        const OpRef &o = ops[in.op];
        const OpSig &sig = table[o.opcode];
        for (size_t k = 0; k < sig.args.size(); k++) {
          if (sig.args[k].is_input) continue;
          uint64_t addr = code[in.code_at + k];
          if (ringEligible(o, sig.args[k])) {
            uint32_t slot = ring.alloc(addr, sig.args[k].comps);
            if (slot != ~0u) out_slot[addr] = slot;
          } else {
            ring.drop(addr);
          }
        }
      }
    // a level that produces more than the ring holds wraps around inside itself: results whose slot was
    // taken again by a later result of the same level are written to the memory image only
    for (auto it = out_slot.begin(); it != out_slot.end();) {
      auto w = ring.where.find(it->first);
      if (w == ring.where.end() || w->second.first != it->second) it = out_slot.erase(it);
      else ++it;
    }
    std::stable_sort(bundles.begin(), bundles.end(), [](const Bundle &a, const Bundle &b) { return a.cost > b.cost; });
    std::fill(load.begin(), load.end(), 0);
    for (auto &bd : bundles) {
      uint32_t w = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
      load[w] += bd.cost;
      auto &s = streams[w];
      for (uint32_t p = 0; p < bd.len; p++) {
      const uint16_t opcode = ops[member(bd, 0, p).op].opcode;
      const OpSig &sig = table[opcode];
      uint32_t na = (uint32_t)sig.args.size();
      uint32_t stride = bd.wide ? 32 : 4;
      const OpRef bo{opcode, bd.wide ? 1u : 8u};
      bool dual = false;
      if (use_cache && !bd.wide)
        for (uint32_t i = 0; i < bd.count && !dual; i++) {
          const Inst &in = member(bd, i, p);
          for (uint32_t k = 0; k < na; k++)
            if (!sig.args[k].is_input && out_slot.count(code[in.code_at + k])) dual = true;
        }
      room(s, 1 + (size_t)na * stride * (dual ? 2 : 1));
      s.push_back(HeaderBits::make(opcode, bd.wide, bd.count, na) | ((dual ? 1u : 0u) << 22));
      size_t base = s.size();
      s.resize(base + (size_t)na * stride * (dual ? 2 : 1), 0xffffffffu);
      for (uint32_t i = 0; i < bd.count; i++) {
        const Inst &in = member(bd, i, p);
        for (uint32_t k = 0; k < na; k++) {
          uint64_t addr = code[in.code_at + k];
          uint32_t image_word = (uint32_t)(addr / 8), word = image_word;
          if (use_gpool && ringEligible(bo, sig.args[k]) && sig.args[k].is_input) {
            auto pc_it = pool_of_const.find(addr);
            if (pc_it != pool_of_const.end()) {
              word = 0x80000000u | (pc_it->second * 8);
              out.n_ring_reads++;
            } else {
              out.n_image_reads++;
            }
          }
          if (use_cache && ringEligible(bo, sig.args[k])) {
            if (sig.args[k].is_input) {
              auto pc_it = pool_of_const.find(addr);
              auto it = ring.where.find(addr);
              if (it != ring.where.end() && it->second.second == sig.args[k].comps && !out_slot.count(addr)) {
                word = 0x80000000u | (it->second.first * 8);
                out.n_ring_reads++;
              } else if (pc_it != pool_of_const.end()) {
                word = 0x80000000u | (out.pool_offset_doubles + pc_it->second * 8);
                out.n_ring_reads++;
              } else {
                out.n_image_reads++;
              }
            } else {
              auto it = out_slot.find(addr);
              if (it != out_slot.end()) {
                word = 0x80000000u | (it->second * 8);
                s[base + (size_t)(na + k) * stride + i] = image_word;  // write-through to the memory image
                out.n_dual_writes++;
              } else {
                out.n_image_writes++;
              }
            }
          }
          s[base + (size_t)k * stride + i] = word;
        }
      }
      // idle sub-slots repeat the first instruction's addresses so that address arithmetic stays in range
      for (uint32_t i = bd.count; i < stride; i++)
        for (uint32_t k = 0; k < na * (dual ? 2u : 1u); k++) s[base + (size_t)k * stride + i] = s[base + (size_t)k * stride];
      }
    }
    for (auto &s : streams) s.push_back(HeaderBits::make(OP_BARRIER, false, 1, 0));
    out.level_bundles.push_back((uint32_t)bundles.size() | (n_coop_level << 16));
  }
  for (auto &s : streams) s.push_back(HeaderBits::make(OP_END, false, 1, 0));
  out.stream_begin.assign(W + 1, 0);
  for (uint32_t w = 0; w < W; w++) {
    out.stream_begin[w] = out.stream.size();
    out.stream.insert(out.stream.end(), streams[w].begin(), streams[w].end());
    // pad so that a one-record look-ahead never reads past the allocation, and to whole pages
    for (int i = 0; i < 16; i++) out.stream.push_back(HeaderBits::make(OP_END, false, 1, 0));
    if (PG)
      while (out.stream.size() & (PG - 1)) out.stream.push_back(HeaderBits::make(OP_END, false, 1, 0));
  }
  out.stream_begin[W] = out.stream.size();

  // ---- 5. ports and constants
  out.inputs = makePortLayout(src.inputs, tape_lanes == 1 ? 8 : tape_lanes, src.memory_size);
  out.parameters = makePortLayout(src.parameters, tape_lanes == 1 ? 8 : tape_lanes, src.memory_size);
  out.outputs = makePortLayout(src.outputs, tape_lanes == 1 ? 8 : tape_lanes, src.memory_size);
  for (auto &p : src.constants) {
    if (p.address % 8 || p.size % 8) throw std::runtime_error("constant not 8-byte aligned");
    if (p.offset + p.size > src.const_data.size()) throw std::runtime_error("constant outside the constant blob");
    if (p.address + p.size > src.memory_size) throw std::runtime_error("constant outside the memory image");
    for (uint32_t e = 0; e < p.size / 8; e++) {
      uint64_t bits;
      std::memcpy(&bits, src.const_data.data() + p.offset + (size_t)e * 8, 8);
      out.const_idx.push_back((uint32_t)(p.address / 8 + e));
      out.const_bits.push_back(bits);
    }
  }
}

}  // namespace trx
