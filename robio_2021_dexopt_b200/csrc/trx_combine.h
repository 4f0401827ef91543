// trx_combine.h -- x_prep interleaved into x_prog (objective lowering, default path of trx_objective_create).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "../../include/trx_file.h"

namespace trx {

// interleaves x_prep into x_prog: each prepare instruction directly after the compute instruction
// whose arguments it repeats (gradients.cpp:42-56 emits them in the same order)
inline void combineForward(const trxfile::Program &prog, const trxfile::Program &prep, trxfile::Program &out) {
  out = trxfile::Program();
  out.lane_width = prog.lane_width;
  out.memory_size = std::max(prog.memory_size, prep.memory_size);
  out.ops = prog.ops;
  std::vector<uint32_t> prep_op(prep.ops.size());
  for (size_t i = 0; i < prep.ops.size(); i++) {
    prep_op[i] = (uint32_t)out.ops.size();
    out.ops.push_back(prep.ops[i]);
  }
  out.inputs = prog.inputs;
  out.parameters = prog.parameters;
  out.outputs = prog.outputs;
  out.constants = prog.constants;
  out.const_data = prog.const_data;
  out.goals = prog.goals;
  uint64_t pa = 0, pb = 0;
  while (pa < prog.code.size()) {
    uint64_t oi = prog.code[pa];
    size_t na = prog.ops[oi].args.size();
    out.code.insert(out.code.end(), prog.code.begin() + pa, prog.code.begin() + pa + 1 + na);
    out.n_instructions++;
    if (pb < prep.code.size()) {
      uint64_t qi = prep.code[pb];
      size_t nq = prep.ops[qi].args.size();
      bool match = prep.ops[qi].name == "prepare_" + prog.ops[oi].name && nq >= na;
      for (size_t k = 0; match && k < na; k++) match = prep.code[pb + 1 + k] == prog.code[pa + 1 + k];
      if (match) {
        out.code.push_back(prep_op[qi]);
        out.code.insert(out.code.end(), prep.code.begin() + pb + 1, prep.code.begin() + pb + 1 + nq);
        out.n_instructions++;
        pb += 1 + nq;
      }
    }
    pa += 1 + na;
  }
  if (pb != prep.code.size()) throw std::runtime_error("x_prep does not mirror x_prog");
}

}  // namespace trx
