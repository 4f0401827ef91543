// trx_fuse.h -- EXPERIMENTAL latency-oriented lowering of the objective (forward+prepare, adjoint).
//
// Not part of the product library: compiled only with -DTRX_WITH_EXPERIMENTS (and by the test-only host
// emulation, which keeps checking it against the reference).  It measured slower than the staged level-scheduled
// kernel (profiles/README.md, round 1) and is kept as a record of that experiment.
//
// lowerProgram (trx_lower.h) keeps the reference's execution contract literally: every slot of
// the tape lives in the memory image and every dependency level ends in a block barrier.  That
// is what a generic drop-in Executable needs, and it is latency-bound (profiles/README.md).
// When the objective owns the whole sequence  x_prog ; x_prep ; x_bprop  (solvers/gd.h:50-63)
// more is known, and this lowering uses it:
//
//   * chain scheduling instead of level scheduling.  Instructions are placed on the W warps by
//     list scheduling in tape order: an instruction goes where its operands were produced
//     unless another warp would finish it earlier even after paying for a block barrier.  A
//     dependency inside one warp needs no barrier (program order + __syncwarp); barriers
//     ("epochs") are only inserted where a value crosses warps.  Dependent chains -- a finger's
//     forward kinematics, one end effector's contact terms -- run back to back on one warp.
//   * an on-chip slot ring per warp (shared memory).  Every lane-typed result is written into
//     the producing warp's ring; a later instruction of the same warp reads it from there as
//     long as it has not been overwritten.  The decision is static: the lowering simulates the
//     ring and encodes "ring offset" or "memory-image slot" in each argument word.
//   * write-through only where needed.  A result is also written to the memory image if some
//     reader cannot take it from the ring (other warp, block operator, evicted) or if it is in
//     the `persist` set (read by a later program, or a port).  Results whose readers are all
//     ring hits never reach HBM.
//   * x_prep is interleaved into x_prog (each prepare right after its compute instruction), so
//     the linearisation state is computed while the operands are still on chip.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <unordered_map>
#include <unordered_set>

#include "trx_combine.h"
#include "trx_lower.h"

namespace trx {

struct FuseOptions {
  uint32_t n_warps = 16;
  uint32_t ring_slots = 192;    // 64-byte slots per warp (192 -> 12 KB per warp, 192 KB per CTA)
  uint32_t sync_penalty = 900;  // cycles charged for taking a value across warps (barrier + skew)
  uint32_t op_base_cost = 140;  // cycles per instruction: dispatch + operand latency
  uint32_t op_comp_cost = 10;   // + cycles per component moved
};


inline void lowerFused(const trxfile::Program &src, const std::unordered_set<uint64_t> *persist, const FuseOptions &opt,
                       LoweredProgram &out) {
  const auto &table = opTable();
  out = LoweredProgram();
  out.memory_size = (src.memory_size + 255) / 256 * 256;
  out.n_warps = opt.n_warps;
  out.ring_doubles = opt.ring_slots * 8;
  const uint32_t W = opt.n_warps;
  if (W < 1 || W > 16) throw std::runtime_error("n_warps out of range (1..16)");
  if (opt.ring_slots * 8 > 0xffff) throw std::runtime_error("ring too large for 16-bit offsets");
  if (const char *e = getenv("TRX_OP_COST")) const_cast<FuseOptions &>(opt).op_base_cost = (uint32_t)atoi(e);
  if (src.memory_size / 8 >= 0x7fffffffull) throw std::runtime_error("memory image too large for 31-bit slot indices");

  // ---- operators
  struct OpRef {
    uint16_t opcode;
    uint32_t lane_width;
  };
  std::vector<OpRef> ops(src.ops.size());
  for (size_t i = 0; i < src.ops.size(); i++) {
    uint16_t opcode;
    uint32_t lw;
    if (!findOp(src.ops[i].name, opcode, lw))
      throw std::runtime_error("unsupported operator on the device engine (no CPU fallback): " + src.ops[i].name);
    const OpSig &sig = table[opcode];
    if (sig.args.size() != src.ops[i].args.size()) throw std::runtime_error("argument count mismatch for " + src.ops[i].name);
    ops[i] = {opcode, lw};
    if (sig.mode == 'F') out.has_forward = true;
    if (sig.mode == 'R') out.has_reverse = true;
  }
  struct Inst {
    uint32_t op;
    uint64_t code_at;
    int32_t warp;  // -1 = cooperative (all warps)
    uint32_t epoch;
    uint32_t local_level;
  };
  std::vector<Inst> insts;
  insts.reserve(src.n_instructions);
  for (uint64_t pc = 0; pc < src.code.size();) {
    uint64_t oi = src.code[pc];
    if (oi >= ops.size()) throw std::runtime_error("invalid op");
    size_t na = table[ops[oi].opcode].args.size();
    insts.push_back({(uint32_t)oi, pc + 1, 0, 0, 0});
    pc += 1 + na;
  }
  out.n_instructions = insts.size();

  auto forEachKey = [&](const Inst &in, size_t k, const std::function<void(uint64_t)> &f) {
    const ArgSig &a = table[ops[in.op].opcode].args[k];
    uint64_t addr = src.code[in.code_at + k];
    if (!isBlock(a.kind)) {
      f(addr);
      return;
    }
    uint32_t eb = blockElemBytes(a.kind, 8), size = src.ops[in.op].args[k].size;
    for (uint32_t off = 0; off < size; off += eb) f(addr + off);
  };
  std::vector<uint64_t> addrs;
  for (auto &in : insts) {
    size_t na = table[ops[in.op].opcode].args.size();
    for (size_t k = 0; k < na; k++) {
      uint64_t a = src.code[in.code_at + k];
      if (a % 8 || a + src.ops[in.op].args[k].size > src.memory_size)
        throw std::runtime_error("argument address misaligned or outside the memory image");
      out.arg_bytes_per_tile += src.ops[in.op].args[k].size;
      forEachKey(in, k, [&](uint64_t key) { addrs.push_back(key); });
    }
  }
  std::sort(addrs.begin(), addrs.end());
  addrs.erase(std::unique(addrs.begin(), addrs.end()), addrs.end());
  auto slotId = [&](uint64_t a) -> size_t {
    return (size_t)(std::lower_bound(addrs.begin(), addrs.end(), a) - addrs.begin());
  };

  // ---- 1. chain scheduling (list scheduling in tape order with a cross-warp penalty)
  struct KeyState {
    int32_t w_warp = -2;  // -2 none, -1 cooperative, >= 0 warp
    uint32_t w_epoch = 0;
    uint64_t w_finish = 0;
    int32_t r_warp = -2;  // -2 none, -3 several warps
    uint32_t r_epoch = 0;
    uint64_t r_finish = 0;
  };
  std::vector<KeyState> keys(addrs.size());
  std::vector<uint64_t> avail(W, 0);
  std::vector<uint32_t> cur_epoch(W, 1);
  struct Pred {
    int32_t warp;
    uint32_t epoch;
    uint64_t finish;
  };
  std::vector<Pred> preds;
  std::vector<std::vector<uint32_t>> lists(W);  // per warp: instruction indices in execution order
  std::vector<uint32_t> coop_list;
  uint32_t max_epoch = 1;
  for (uint32_t i = 0; i < insts.size(); i++) {
    Inst &in = insts[i];
    const OpSig &sig = table[ops[in.op].opcode];
    size_t na = sig.args.size();
    preds.clear();
    for (size_t k = 0; k < na; k++) {
      bool is_in = sig.args[k].is_input;
      forEachKey(in, k, [&](uint64_t key) {
        const KeyState &ks = keys[slotId(key)];
        if (ks.w_warp != -2) preds.push_back({ks.w_warp, ks.w_epoch, ks.w_finish});
        if (!is_in && ks.r_warp != -2) preds.push_back({ks.r_warp, ks.r_epoch, ks.r_finish});
      });
    }
    uint64_t cost = opt.op_base_cost + (uint64_t)opt.op_comp_cost * sig.cost;
    uint64_t finish;
    if (sig.coop) {
      uint32_t e = 1;
      for (uint32_t w = 0; w < W; w++) e = std::max(e, cur_epoch[w]);
      for (auto &p : preds) e = std::max(e, p.epoch + 1);
      uint64_t start = 0;
      for (uint32_t w = 0; w < W; w++) start = std::max(start, avail[w]);
      for (auto &p : preds) start = std::max(start, p.finish + opt.sync_penalty);
      uint64_t bytes = 0;
      for (size_t k = 0; k < na; k++) bytes += src.ops[in.op].args[k].size;
      finish = start + 400 + bytes / 16;
      in.warp = -1;
      in.epoch = e;
      for (uint32_t w = 0; w < W; w++) {
        cur_epoch[w] = e;
        avail[w] = finish;
      }
      coop_list.push_back(i);
    } else {
      // candidate warps: those holding a predecessor, and the one that is free first
      uint32_t best_w = 0;
      uint64_t best_est = ~0ull;
      auto consider = [&](uint32_t w) {
        uint64_t est = avail[w];
        for (auto &p : preds) est = std::max(est, p.finish + ((p.warp == (int32_t)w) ? 0 : opt.sync_penalty));
        if (est < best_est || (est == best_est && avail[w] < avail[best_w])) {
          best_est = est;
          best_w = w;
        }
      };
      uint32_t idle = (uint32_t)(std::min_element(avail.begin(), avail.end()) - avail.begin());
      consider(idle);
      for (auto &p : preds)
        if (p.warp >= 0) consider((uint32_t)p.warp);
      uint32_t e = cur_epoch[best_w];
      for (auto &p : preds) e = std::max(e, (p.warp == (int32_t)best_w) ? p.epoch : p.epoch + 1);
      in.warp = (int32_t)best_w;
      in.epoch = e;
      cur_epoch[best_w] = e;
      finish = best_est + cost;
      avail[best_w] = finish;
      lists[best_w].push_back(i);
    }
    max_epoch = std::max(max_epoch, in.epoch);
    for (size_t k = 0; k < na; k++) {
      bool is_in = sig.args[k].is_input;
      forEachKey(in, k, [&](uint64_t key) {
        KeyState &ks = keys[slotId(key)];
        if (is_in) {
          if (ks.r_warp == -2) ks.r_warp = in.warp;
          else if (ks.r_warp != in.warp) ks.r_warp = -3;
          ks.r_epoch = std::max(ks.r_epoch, in.epoch);
          ks.r_finish = std::max(ks.r_finish, finish);
        } else {
          ks.w_warp = in.warp;
          ks.w_epoch = in.epoch;
          ks.w_finish = finish;
          ks.r_warp = -2;
          ks.r_epoch = 0;
          ks.r_finish = 0;
        }
      });
    }
  }
  out.n_levels = max_epoch;

  // ---- 2. per (warp, epoch): local levels, then same-opcode bundles
  struct Bundle {
    uint16_t opcode;
    bool wide;
    uint32_t epoch;
    std::vector<uint32_t> members;  // instruction indices (<= 4 or <= 32)
  };
  std::vector<std::vector<Bundle>> bundles(W);
  {
    std::vector<uint32_t> stamp_w(addrs.size(), 0), stamp_r(addrs.size(), 0), lvl_w(addrs.size(), 0), lvl_r(addrs.size(), 0);
    uint32_t stamp = 0;
    std::vector<uint32_t> seg;
    for (uint32_t w = 0; w < W; w++) {
      auto &list = lists[w];
      size_t pos = 0;
      while (pos < list.size()) {
        uint32_t e = insts[list[pos]].epoch;
        seg.clear();
        while (pos < list.size() && insts[list[pos]].epoch == e) seg.push_back(list[pos++]);
        stamp++;
        for (uint32_t i : seg) {
          Inst &in = insts[i];
          const OpSig &sig = table[ops[in.op].opcode];
          uint32_t lvl = 0;
          for (size_t k = 0; k < sig.args.size(); k++) {
            size_t id = slotId(src.code[in.code_at + k]);
            if (stamp_w[id] == stamp) lvl = std::max(lvl, lvl_w[id]);
            if (!sig.args[k].is_input && stamp_r[id] == stamp) lvl = std::max(lvl, lvl_r[id]);
          }
          in.local_level = lvl + 1;
          for (size_t k = 0; k < sig.args.size(); k++) {
            size_t id = slotId(src.code[in.code_at + k]);
            if (sig.args[k].is_input) {
              if (stamp_r[id] != stamp) {
                stamp_r[id] = stamp;
                lvl_r[id] = 0;
              }
              lvl_r[id] = std::max(lvl_r[id], in.local_level);
            } else {
              stamp_w[id] = stamp;
              lvl_w[id] = in.local_level;
              stamp_r[id] = stamp;
              lvl_r[id] = in.local_level;
            }
          }
        }
        std::stable_sort(seg.begin(), seg.end(), [&](uint32_t a, uint32_t b) {
          const Inst &ia = insts[a], &ib = insts[b];
          if (ia.local_level != ib.local_level) return ia.local_level < ib.local_level;
          const OpRef &oa = ops[ia.op], &ob = ops[ib.op];
          if (oa.opcode != ob.opcode) return oa.opcode < ob.opcode;
          return oa.lane_width < ob.lane_width;
        });
        for (size_t a = 0; a < seg.size();) {
          const Inst &ia = insts[seg[a]];
          const OpRef &oa = ops[ia.op];
          bool wide = oa.lane_width == 1;
          size_t cap = wide ? 32 : 4, b = a;
          Bundle bd{oa.opcode, wide, e, {}};
          while (b < seg.size() && b - a < cap) {
            const Inst &ib = insts[seg[b]];
            if (ib.local_level != ia.local_level || ops[ib.op].opcode != oa.opcode || ops[ib.op].lane_width != oa.lane_width)
              break;
            bd.members.push_back(seg[b]);
            b++;
          }
          bundles[w].push_back(std::move(bd));
          out.n_bundles++;
          a = b;
        }
      }
    }
  }

  // ---- 3. ring residency (static): which reads hit the producing warp's ring, which results must
  //         also reach the memory image
  struct RingSim {
    std::vector<int64_t> owner;  // per slot: value id (address) or -1
    std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> where;  // address -> (slot offset, comps)
    uint32_t head = 0;
    explicit RingSim(uint32_t n) : owner(n, -1) {}
    uint32_t alloc(uint64_t addr, uint32_t comps) {
      uint32_t n = (uint32_t)owner.size();
      if (comps > n) return ~0u;
      if (head + comps > n) head = 0;
      for (uint32_t s = head; s < head + comps; s++) {
        if (owner[s] >= 0) {
          auto it = where.find((uint64_t)owner[s]);
          if (it != where.end()) {
            for (uint32_t t = it->second.first; t < it->second.first + it->second.second; t++) owner[t] = -1;
            where.erase(it);
          }
        }
      }
      auto old = where.find(addr);  // a slot that is written again (gradient accumulation)
      if (old != where.end()) {
        for (uint32_t t = old->second.first; t < old->second.first + old->second.second; t++) owner[t] = -1;
        where.erase(old);
      }
      for (uint32_t s = head; s < head + comps; s++) owner[s] = (int64_t)addr;
      where[addr] = {head, comps};
      uint32_t r = head;
      head += comps;
      return r;
    }
  };
  auto ringEligible = [&](const OpRef &o, const ArgSig &a) { return o.lane_width == 8 && a.kind == ARG_T && a.comps > 0; };
  // produced-in-this-program values: address -> needs the memory image / has ring readers
  std::unordered_set<uint64_t> need_image, ring_read;
  // a value written by an instruction whose result is not ring-eligible is in the image anyway
  for (int phase = 0; phase < 2; phase++) {
    std::vector<std::vector<uint32_t>> streams(W);
    std::vector<size_t> next_bundle(W, 0);
    std::vector<RingSim> rings;
    for (uint32_t w = 0; w < W; w++) rings.emplace_back(opt.ring_slots);
    size_t next_coop = 0;
    for (uint32_t e = 1; e <= max_epoch; e++) {
      // cooperative records of this epoch first, identical in every stream
      while (next_coop < coop_list.size() && insts[coop_list[next_coop]].epoch == e) {
        const Inst &in = insts[coop_list[next_coop++]];
        const OpRef &o = ops[in.op];
        const OpSig &sig = table[o.opcode];
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
        if (phase == 0) {
          for (uint32_t k = 0; k < na; k++)
            forEachKey(in, k, [&](uint64_t key) {
              need_image.insert(key);
              // block operators write the image: ring copies of those slots are stale from now on
              if (!sig.args[k].is_input)
                for (auto &r : rings) {
                  auto it = r.where.find(key);
                  if (it != r.where.end()) {
                    for (uint32_t t = it->second.first; t < it->second.first + it->second.second; t++) r.owner[t] = -1;
                    r.where.erase(it);
                  }
                }
            });
        } else {
          for (uint32_t k = 0; k < na; k++)
            if (!sig.args[k].is_input)
              forEachKey(in, k, [&](uint64_t key) {
                for (auto &r : rings) {
                  auto it = r.where.find(key);
                  if (it != r.where.end()) {
                    for (uint32_t t = it->second.first; t < it->second.first + it->second.second; t++) r.owner[t] = -1;
                    r.where.erase(it);
                  }
                }
              });
          for (auto &s : streams) {
            s.push_back(HeaderBits::make(o.opcode, false, 1, na, true));
            for (uint32_t k = 0; k < na; k++) s.push_back((uint32_t)(src.code[in.code_at + k] / 8));
            s.push_back(imm0);
            s.push_back(imm1);
          }
        }
      }
      for (uint32_t w = 0; w < W; w++) {
        RingSim &ring = rings[w];
        auto &s = streams[w];
        while (next_bundle[w] < bundles[w].size() && bundles[w][next_bundle[w]].epoch == e) {
          const Bundle &bd = bundles[w][next_bundle[w]++];
          const OpSig &sig = table[bd.opcode];
          const OpRef o{bd.opcode, bd.wide ? 1u : 8u};
          uint32_t na = (uint32_t)sig.args.size();
          uint32_t stride = bd.wide ? 32 : 4;
          uint32_t count = (uint32_t)bd.members.size();
          // words[k][i] primary, words2[k][i] secondary (image copy of an output that also lives in the ring)
          std::vector<uint32_t> words((size_t)na * stride, 0), words2((size_t)na * stride, 0xffffffffu);
          bool dual = false;
          // outputs first: ring allocation (evicts), then inputs are resolved against what survived
          std::vector<uint32_t> ring_off((size_t)na * count, ~0u);
          for (uint32_t i = 0; i < count; i++) {
            const Inst &in = insts[bd.members[i]];
            for (uint32_t k = 0; k < na; k++) {
              if (sig.args[k].is_input) continue;
              uint64_t addr = src.code[in.code_at + k];
              // a slot that is written again (gradient accumulation): copies in other warps' rings are stale
              for (uint32_t w2 = 0; w2 < W; w2++) {
                if (w2 == w) continue;
                auto it = rings[w2].where.find(addr);
                if (it != rings[w2].where.end()) {
                  for (uint32_t t = it->second.first; t < it->second.first + it->second.second; t++) rings[w2].owner[t] = -1;
                  rings[w2].where.erase(it);
                }
              }
              if (!ringEligible(o, sig.args[k])) {
                auto it = ring.where.find(addr);
                if (it != ring.where.end()) {
                  for (uint32_t t = it->second.first; t < it->second.first + it->second.second; t++) ring.owner[t] = -1;
                  ring.where.erase(it);
                }
                continue;
              }
              ring_off[(size_t)k * count + i] = ring.alloc(addr, sig.args[k].comps);
            }
          }
          for (uint32_t i = 0; i < count; i++) {
            const Inst &in = insts[bd.members[i]];
            for (uint32_t k = 0; k < na; k++) {
              uint64_t addr = src.code[in.code_at + k];
              uint32_t image_word = (uint32_t)(addr / 8);
              uint32_t word = image_word;
              if (sig.args[k].is_input) {
                bool hit = false;
                if (ringEligible(o, sig.args[k])) {
                  auto it = ring.where.find(addr);
                  if (it != ring.where.end() && it->second.second == sig.args[k].comps) {
                    hit = true;
                    word = 0x80000000u | (it->second.first * 8);
                    if (phase == 0) ring_read.insert(addr);
                    out.n_ring_reads += phase;
                  }
                }
                if (!hit) {
                  if (phase == 0) need_image.insert(addr);
                  out.n_image_reads += phase;
                }
              } else {
                uint32_t ro = ring_off[(size_t)k * count + i];
                bool in_ring = ro != ~0u;
                if (phase == 1) {
                  bool to_image = !in_ring || need_image.count(addr) || (persist ? persist->count(addr) != 0 : true);
                  bool to_ring = in_ring && ring_read.count(addr);
                  if (to_ring && to_image) {
                    word = 0x80000000u | (ro * 8);
                    words2[(size_t)k * stride + i] = image_word;
                    dual = true;
                    out.n_dual_writes++;
                  } else if (to_ring) {
                    word = 0x80000000u | (ro * 8);
                    out.n_ring_only_writes++;
                  } else if (to_image) {
                    word = image_word;
                    out.n_image_writes++;
                  } else {
                    // nobody reads it: keep the store on chip
                    word = in_ring ? (0x80000000u | (ro * 8)) : image_word;
                    out.n_ring_only_writes++;
                  }
                }
              }
              words[(size_t)k * stride + i] = word;
            }
          }
          if (phase == 1) {
            for (uint32_t i = count; i < stride; i++)
              for (uint32_t k = 0; k < na; k++) {
                words[(size_t)k * stride + i] = words[(size_t)k * stride];
                words2[(size_t)k * stride + i] = words2[(size_t)k * stride];
              }
            s.push_back(HeaderBits::make(bd.opcode, bd.wide, count, na) | ((dual ? 1u : 0u) << 22));
            s.insert(s.end(), words.begin(), words.end());
            if (dual) s.insert(s.end(), words2.begin(), words2.end());
          }
        }
        if (phase == 1 && e < max_epoch) s.push_back(HeaderBits::make(OP_BARRIER, false, 1, 0));
      }
    }
    if (phase == 1) {
      for (auto &s : streams) s.push_back(HeaderBits::make(OP_END, false, 1, 0));
      out.stream_begin.assign(W + 1, 0);
      for (uint32_t w = 0; w < W; w++) {
        out.stream_begin[w] = out.stream.size();
        out.stream.insert(out.stream.end(), streams[w].begin(), streams[w].end());
        for (int i = 0; i < 128; i++) out.stream.push_back(HeaderBits::make(OP_END, false, 1, 0));
      }
      out.stream_begin[W] = out.stream.size();
    }
  }

  // ---- 4. ports and constants (as lowerProgram)
  out.inputs = makePortLayout(src.inputs, 8, src.memory_size);
  out.parameters = makePortLayout(src.parameters, 8, src.memory_size);
  out.outputs = makePortLayout(src.outputs, 8, src.memory_size);
  for (auto &p : src.constants) {
    if (p.address % 8 || p.size % 8) throw std::runtime_error("constant not 8-byte aligned");
    if (p.offset + p.size > src.const_data.size()) throw std::runtime_error("constant outside the constant blob");
    for (uint32_t e = 0; e < p.size / 8; e++) {
      uint64_t bits;
      std::memcpy(&bits, src.const_data.data() + p.offset + (size_t)e * 8, 8);
      out.const_idx.push_back((uint32_t)(p.address / 8 + e));
      out.const_bits.push_back(bits);
    }
  }
}

// The objective's two sweeps: forward = x_prog with x_prep interleaved, adjoint = x_bprop.  Only what the
// adjoint reads (prepare state, operands of operators without a prepare variant) and the ports are
// written through to the memory image.
inline void lowerObjective(const trxfile::Program &prog, const trxfile::Program &prep, const trxfile::Program &bprop,
                           const FuseOptions &opt, LoweredProgram &fwd, LoweredProgram &adj) {
  trxfile::Program combined;
  combineForward(prog, prep, combined);
  std::unordered_set<uint64_t> persist;
  const auto &table = opTable();
  for (uint64_t pc = 0; pc < bprop.code.size();) {
    const auto &op = bprop.ops[bprop.code[pc]];
    uint16_t opcode;
    uint32_t lw;
    if (!findOp(op.name, opcode, lw))
      throw std::runtime_error("unsupported operator on the device engine (no CPU fallback): " + op.name);
    const OpSig &sig = table[opcode];
    for (size_t k = 0; k < op.args.size(); k++) {
      if (!op.args[k].is_input) continue;
      uint64_t addr = bprop.code[pc + 1 + k];
      if (isBlock(sig.args[k].kind)) {
        uint32_t eb = blockElemBytes(sig.args[k].kind, 8);
        for (uint32_t off = 0; off < op.args[k].size; off += eb) persist.insert(addr + off);
      } else {
        persist.insert(addr);
      }
    }
    pc += 1 + op.args.size();
  }
  for (auto &p : prog.outputs) persist.insert(p.address);
  combined.memory_size = std::max(combined.memory_size, bprop.memory_size);
  lowerFused(combined, &persist, opt, fwd);
  std::unordered_set<uint64_t> persist_adj;
  for (auto &p : bprop.outputs) persist_adj.insert(p.address);
  lowerFused(bprop, &persist_adj, opt, adj);
}

}  // namespace trx
