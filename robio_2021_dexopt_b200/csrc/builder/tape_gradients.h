// tape_gradients.h -- reverse/forward-mode program derivation by tape rewriting.
//
// Product-side restatement of tractor's buildGradients
// (tractor/src/core/gradients.cpp:22-407) on the neutral program form, so that the
// product can differentiate tapes it built itself.  Given a forward program it derives
//   prep   per-instruction linearisation state            gradients.cpp:35-65
//   bprop  the adjoint sweep, g_in = J^T g_out             gradients.cpp:67-234
//          (reverse variants in reverse order, fan-out accumulation through
//          temp + move/add pairs :120-179, zero prologue :205-233)
//   fprop  the tangent sweep                               gradients.cpp:236-316
//   hprop  fprop followed by bprop                         gradients.cpp:344-362
//   accu   x (+) delta per input                           gradients.cpp:364-406
// Addressing follows the reference: prep state is allocated above the forward image,
// the gradient slot of forward address a is a + prep.memory_size.
// Parity is checked program-for-program against the reference's own output in
// tests/test_builder.py.
#pragma once
#include "tape_passes.h"

namespace trxb {

struct Alloc {
  uint64_t top = 32;  // allocator.h:19
  uint64_t alloc(uint32_t size, uint32_t lanes) {
    uint64_t align = (lanes > 1) ? 32 : 8;  // alignas(32) Batch (batch.h:42), natural alignment for double
    while (top % align) top++;
    uint64_t r = top;
    top += size;
    return r;
  }
};

struct GradientPrograms {
  trxfile::Program prep, fprop, bprop, hprop, accu;
};

inline std::string variantName(const std::string &name, const char *mode) { return std::string(mode) + "_" + name; }
inline bool hasOp(const std::string &name) {
  uint16_t opcode;
  uint32_t lw;
  return trx::findOp(name, opcode, lw);
}

inline std::string typedOp(const char *base, const trxfile::ArgDesc &a) {
  // compute-mode move/add/zero operator for a value with this layout (Operator::find<compute, op_move>(type))
  uint32_t comps = a.size / (8 * a.lanes);
  const char *group = comps == 1 ? "" : comps == 3 ? "_fg_vec3" : comps == 4 ? "_fg_quat" : comps == 6 ? "_fg_twist"
                      : comps == 7 ? "_fg_pose" : comps == 9 ? "_fg_mat3" : nullptr;
  if (!group) throw std::runtime_error("no typed operator for a value of " + std::to_string(comps) + " components");
  return std::string(base) + group + (a.lanes == 8 ? "_8d" : "_d");
}

inline void packOffsets(std::vector<trxfile::Port> &ports) {
  uint64_t off = 0;
  for (auto &p : ports) {
    p.offset = off;
    off += p.size;
  }
}

// want_tangent = false skips fprop / hprop
inline void buildGradients(const trxfile::Program &src, GradientPrograms &out, bool want_tangent = true) {
  auto insts = decode(src);
  out = GradientPrograms();
  trxfile::Program &prep = out.prep, &fprop = out.fprop, &bprop = out.bprop, &hprop = out.hprop, &accu = out.accu;
  for (auto *p : {&prep, &fprop, &bprop, &hprop, &accu}) p->lane_width = src.lane_width;

  // ---- prep (gradients.cpp:35-65)
  std::vector<int64_t> prep_at(insts.size(), -1);  // position of the prepare instruction's first arg in prep.code
  std::vector<uint32_t> prep_nargs(insts.size(), 0);
  {
    Alloc alloc;
    alloc.top = src.memory_size;
    auto index = opIndex(prep);
    for (size_t i = 0; i < insts.size(); i++) {
      const auto &in = insts[i];
      std::string pname = variantName(src.ops[in.op].name, "prepare");
      if (!hasOp(pname)) continue;
      uint32_t id = internOp(prep, index, pname);
      const auto &pop = prep.ops[id];
      prep.code.push_back(id);
      prep_at[i] = (int64_t)prep.code.size();
      prep_nargs[i] = (uint32_t)pop.args.size();
      for (uint32_t k = 0; k < in.n_args; k++) prep.code.push_back(src.code[in.code_at + k]);
      for (uint32_t k = in.n_args; k < pop.args.size(); k++) prep.code.push_back(alloc.alloc(pop.args[k].size, pop.args[k].lanes));
      prep.n_instructions++;
    }
    prep.memory_size = alloc.top;
  }
  const uint64_t G = prep.memory_size;  // gradient slot offset

  // The reference converts every port to its gradient type (Pose -> Twist, Quaternion -> Vector3,
  // gradients.cpp:72-86); here ports are copied with their value layout, which is the same thing for scalars and
  // vectors -- the only port types the product's recorder creates.  Refuse the others instead of mis-sizing them.
  for (auto *ports : {&src.inputs, &src.outputs})
    for (auto &p : *ports) {
      const uint32_t comps = p.size / (8 * (p.lanes ? p.lanes : 1));
      if (comps == 7 || comps == 4)
        throw std::runtime_error("pose- or quaternion-typed input / output ports are not supported by this gradient builder");
    }
  // ---- bprop (gradients.cpp:67-234)
  {
    for (auto p : src.inputs) {
      p.address += G;
      bprop.outputs.push_back(p);
    }
    packOffsets(bprop.outputs);
    for (auto p : src.outputs) {
      p.address += G;
      bprop.inputs.push_back(p);
    }
    packOffsets(bprop.inputs);
    auto index = opIndex(bprop);
    auto src_meta = opMeta(src);
    struct RInst {
      uint32_t op;
      std::vector<uint64_t> args;
    };
    std::vector<RInst> rev;
    rev.reserve(insts.size());
    for (size_t i = insts.size(); i-- > 0;) {
      const auto &in = insts[i];
      std::string rname = variantName(src.ops[in.op].name, "reverse");
      if (!hasOp(rname)) throw std::runtime_error("variant not found: reverse " + src.ops[in.op].name);
      RInst r;
      if (src_meta[in.op].sig->coop) {
        // block operator: the reverse variant takes the original blocks, then their gradient blocks
        std::vector<uint32_t> sizes;
        for (int rep = 0; rep < 2; rep++)
          for (uint32_t k = 0; k < in.n_args; k++) sizes.push_back(src.ops[in.op].args[k].size);
        r.op = internBlockOp(bprop, index, rname, sizes);
      } else {
        r.op = internOp(bprop, index, rname);
      }
      if (prep_at[i] < 0) {
        for (uint32_t k = 0; k < in.n_args; k++) r.args.push_back(src.code[in.code_at + k]);
      } else {
        for (uint32_t k = in.n_args; k < prep_nargs[i]; k++) r.args.push_back(prep.code[prep_at[i] + k]);
      }
      for (uint32_t k = 0; k < in.n_args; k++) r.args.push_back(src.code[in.code_at + k] + G);
      if (r.args.size() != bprop.ops[r.op].args.size())
        throw std::runtime_error("reverse signature mismatch for " + src.ops[in.op].name);
      rev.push_back(std::move(r));
    }
    // fan-out accumulation (gradients.cpp:120-179)
    Alloc alloc;
    alloc.top = src.memory_size + prep.memory_size;
    std::unordered_set<uint64_t> written;
    std::vector<RInst> acc;
    acc.reserve(rev.size() * 2);
    std::vector<RInst> block_zeros;
    for (auto &r : rev) {
      const auto op = bprop.ops[r.op];  // copy: internOp below may grow the table
      uint16_t opcode;
      uint32_t lw;
      trx::findOp(op.name, opcode, lw);
      const trx::OpSig &rsig = trx::opTable()[opcode];
      RInst main = r;
      std::vector<RInst> sums;
      for (size_t k = 0; k < r.args.size(); k++) {
        if (op.args[k].is_input) continue;
        if (trx::isBlock(rsig.args[k].kind)) {
          uint32_t eb = trx::blockElemBytes(rsig.args[k].kind, 8);
          if (rsig.args[k].kind == trx::ARG_SA) {
            // accumulated in place by the operator; zero it once before its first use, and make every
            // later element-wise writer accumulate instead of overwrite
            bool first = true;
            uint32_t n_written = 0, n_elems = 0;
            for (uint32_t off = 0; off < op.args[k].size; off += eb, n_elems++)
              if (written.count(r.args[k] + off)) first = false, n_written++;
            // all or nothing: a block some of whose elements already have a writer would keep stale values in the
            // others (the zero prologue below skips outputs)
            if (n_written != 0 && n_written != n_elems)
              throw std::runtime_error("accumulated gradient block partly written before its first block use");
            if (first) {
              RInst z;
              z.op = internBlockOp(bprop, index, "x_zero_block_d", {op.args[k].size});
              z.args = {r.args[k]};
              block_zeros.push_back(z);
            }
          } else {
            for (uint32_t off = 0; off < op.args[k].size; off += eb)
              if (written.count(r.args[k] + off))
                throw std::runtime_error("block gradient written twice: fan-out of a dense-layer input block");
          }
          for (uint32_t off = 0; off < op.args[k].size; off += eb) written.insert(r.args[k] + off);
          continue;
        }
        if (written.count(r.args[k])) {
          uint64_t a = alloc.alloc(op.args[k].size, op.args[k].lanes);
          uint64_t b = alloc.alloc(op.args[k].size, op.args[k].lanes);
          RInst mv;
          mv.op = internOp(bprop, index, typedOp("move", op.args[k]));
          mv.args = {r.args[k], b};
          RInst ad;
          ad.op = internOp(bprop, index, typedOp("add", op.args[k]));
          ad.args = {b, a, r.args[k]};
          sums.push_back(mv);
          sums.push_back(ad);
          main.args[k] = a;
        }
        written.insert(r.args[k]);
      }
      acc.push_back(std::move(main));
      for (auto &s : sums) acc.push_back(std::move(s));
    }
    bprop.memory_size = alloc.top;
    // zero prologue for gradient slots that are read before they are written (gradients.cpp:205-233)
    std::unordered_set<uint64_t> seen;
    for (auto &p : bprop.inputs) seen.insert(p.address);
    for (auto &z : block_zeros)
      for (uint32_t off = 0; off < bprop.ops[z.op].args[0].size; off += 8) seen.insert(z.args[0] + off);
    std::vector<RInst> zeros = block_zeros;
    for (auto &r : acc) {
      const auto op = bprop.ops[r.op];  // copy: internOp below may grow the table
      uint16_t opcode;
      uint32_t lw;
      trx::findOp(op.name, opcode, lw);
      const trx::OpSig &rsig = trx::opTable()[opcode];
      for (size_t k = 0; k < r.args.size(); k++) {
        if (r.args[k] < G) continue;
        if (trx::isBlock(rsig.args[k].kind)) {
          // element-wise: an element of a gradient block that nothing wrote before it is read is zero
          uint32_t eb = trx::blockElemBytes(rsig.args[k].kind, 8);
          for (uint32_t off = 0; off < op.args[k].size; off += eb)
            if (seen.insert(r.args[k] + off).second && op.args[k].is_input) {
              trxfile::ArgDesc ad;
              ad.size = eb;
              ad.lanes = (rsig.args[k].kind == trx::ARG_TB) ? 8 : 1;
              RInst z;
              z.op = internOp(bprop, index, typedOp("zero", ad));
              z.args = {r.args[k] + off};
              zeros.push_back(z);
            }
          continue;
        }
        if (seen.insert(r.args[k]).second && op.args[k].is_input) {
          RInst z;
          trxfile::ArgDesc ad = op.args[k];
          z.op = internOp(bprop, index, typedOp("zero", ad));
          z.args = {r.args[k]};
          zeros.push_back(z);
        }
      }
    }
    for (auto *list : {&zeros, &acc})
      for (auto &r : *list) {
        bprop.code.push_back(r.op);
        for (auto a : r.args) bprop.code.push_back(a);
        bprop.n_instructions++;
      }
  }

  // ---- fprop (gradients.cpp:236-316)
  if (want_tangent) {
    for (auto p : src.inputs) {
      p.address += G;
      fprop.inputs.push_back(p);
    }
    packOffsets(fprop.inputs);
    for (auto p : src.outputs) {
      p.address += G;
      fprop.outputs.push_back(p);
    }
    packOffsets(fprop.outputs);
    fprop.memory_size = src.memory_size + prep.memory_size;
    auto index = opIndex(fprop);
    auto zero = [&](const trxfile::Port &p) {
      trxfile::ArgDesc ad;
      ad.size = p.size;
      ad.lanes = p.lanes;
      fprop.code.push_back(internOp(fprop, index, typedOp("zero", ad)));
      fprop.code.push_back(p.address + G);
      fprop.n_instructions++;
    };
    for (auto &p : src.constants) zero(p);
    for (auto &p : src.parameters) zero(p);
    for (size_t i = 0; i < insts.size(); i++) {
      const auto &in = insts[i];
      std::string fname = variantName(src.ops[in.op].name, "forward");
      if (!hasOp(fname)) throw std::runtime_error("variant not found: forward " + src.ops[in.op].name);
      uint32_t id;
      {
        uint16_t fopcode;
        uint32_t flw;
        trx::findOp(fname, fopcode, flw);
        if (trx::opTable()[fopcode].coop) {
          // block operator: argument sizes come from the instruction's own blocks (values, then tangents)
          std::vector<uint32_t> sizes;
          for (int rep = 0; rep < 2; rep++)
            for (uint32_t k = 0; k < in.n_args; k++) sizes.push_back(src.ops[in.op].args[k].size);
          id = internBlockOp(fprop, index, fname, sizes);
        } else {
          id = internOp(fprop, index, fname);
        }
      }
      fprop.code.push_back(id);
      size_t n0 = fprop.code.size();
      if (prep_at[i] < 0) {
        for (uint32_t k = 0; k < in.n_args; k++) fprop.code.push_back(src.code[in.code_at + k]);
      } else {
        for (uint32_t k = in.n_args; k < prep_nargs[i]; k++) fprop.code.push_back(prep.code[prep_at[i] + k]);
      }
      for (uint32_t k = 0; k < in.n_args; k++) fprop.code.push_back(src.code[in.code_at + k] + G);
      if (fprop.code.size() - n0 != fprop.ops[id].args.size())
        throw std::runtime_error("forward signature mismatch for " + src.ops[in.op].name);
      fprop.n_instructions++;
    }
  }

  // ---- hprop = fprop ; bprop on one memory (gradients.cpp:344-362)
  if (want_tangent) {
    hprop.inputs = fprop.inputs;
    hprop.outputs = bprop.outputs;
    auto index = opIndex(hprop);
    for (auto *part : {&fprop, &bprop}) {
      for (auto &in : decode(*part)) {
        const auto &pop = part->ops[in.op];
        uint16_t popcode;
        uint32_t plw;
        trx::findOp(pop.name, popcode, plw);
        uint32_t id;
        if (trx::opTable()[popcode].coop) {
          // block operators carry their block sizes in the operator table entry
          std::vector<uint32_t> sizes;
          for (auto &a : pop.args) sizes.push_back(a.size);
          id = internBlockOp(hprop, index, pop.name, sizes);
        } else {
          id = internOp(hprop, index, pop.name);
        }
        hprop.code.push_back(id);
        for (uint32_t k = 0; k < in.n_args; k++) hprop.code.push_back(part->code[in.code_at + k]);
        hprop.n_instructions++;
      }
    }
    hprop.memory_size = std::max(bprop.memory_size, fprop.memory_size);
  }

  // ---- accu: x (+) delta (gradients.cpp:364-406)
  {
    Alloc alloc;
    std::vector<trxfile::Port> aa, bb, xx;
    for (auto p : src.inputs) {
      p.address = alloc.alloc(p.size, p.lanes);
      aa.push_back(p);
      accu.inputs.push_back(p);
    }
    for (auto p : src.inputs) {  // the tangent ports (fprop.inputs) have the layout of the inputs
      p.address = alloc.alloc(p.size, p.lanes);
      bb.push_back(p);
      accu.inputs.push_back(p);
    }
    packOffsets(accu.inputs);
    for (auto p : src.inputs) {
      p.address = alloc.alloc(p.size, p.lanes);
      xx.push_back(p);
      accu.outputs.push_back(p);
    }
    packOffsets(accu.outputs);
    auto index = opIndex(accu);
    for (size_t i = 0; i < aa.size(); i++) {
      trxfile::ArgDesc ad;
      ad.size = aa[i].size;
      ad.lanes = aa[i].lanes;
      uint32_t comps = ad.size / (8 * ad.lanes);
      // Operator::find<compute, op_add>(aa.type(), bb.type()): scalar add, or the manifold plus for
      // orientations / poses (add_fg_quat_vec3, add_fg_pose_twist)
      std::string name = comps == 4 ? std::string("add_fg_quat_vec3") + (ad.lanes == 8 ? "_8d" : "_d")
                         : comps == 7 ? std::string("add_fg_pose_twist") + (ad.lanes == 8 ? "_8d" : "_d")
                                      : typedOp("add", ad);
      accu.code.push_back(internOp(accu, index, name));
      accu.code.push_back(aa[i].address);
      accu.code.push_back(bb[i].address);
      accu.code.push_back(xx[i].address);
      accu.n_instructions++;
    }
    accu.memory_size = alloc.top;
  }
}

}  // namespace trxb
