// ref_common.h -- glue between the reference's own tape library (compiled into
// oracle/_ref/libtractor_ref.so from /root/reference, see Makefile) and the
// neutral TRXB bundle format (include/trx_file.h).
//
// TEST INFRASTRUCTURE ONLY: nothing here is linked into the product library.
#pragma once

#include <cxxabi.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <tractor/core/engine.h>
#include <tractor/core/gradients.h>
#include <tractor/core/operator.h>
#include <tractor/core/ops.h>
#include <tractor/core/program.h>
#include <tractor/core/recorder.h>
#include <tractor/core/var.h>
#include <tractor/core/constraints.h>
#include <tractor/engines/loop.h>
#include <tractor/engines/simple.h>
#include <tractor/geometry/ops.h>

#include "dexopt_ops_ref.h"

#include "../../include/trx_file.h"

namespace refc {

inline std::string demangle(const char *n) {
  int status = 0;
  char *d = abi::__cxa_demangle(n, nullptr, nullptr, &status);
  std::string r = (status == 0 && d) ? d : n;
  free(d);
  return r;
}

// lanes / element kind of a reference TypeInfo, from its demangled C++ name
inline void classify(const tractor::TypeInfo &t, uint8_t &lanes, uint8_t &elem) {
  std::string n = demangle(t.name());
  lanes = 1;
  elem = trxfile::ELEM_F64;
  auto pos = n.find("tractor::Batch<");
  if (pos != std::string::npos) {
    auto comma = n.find(',', pos);
    lanes = (uint8_t)std::atoi(n.c_str() + comma + 1);
    if (n.find("Batch<double") == std::string::npos) elem = trxfile::ELEM_OTHER;
  } else if (n == "unsigned long" || n.find("unsigned long") != std::string::npos) {
    elem = trxfile::ELEM_U64;
  } else if (n.find("float") != std::string::npos) {
    elem = trxfile::ELEM_OTHER;
  }
}

template <class P> trxfile::Port convertPort(const P &port) {
  trxfile::Port q;
  q.address = port.address();
  q.offset = port.offset();
  q.size = (uint32_t)port.size();
  classify(port.typeInfo(), q.lanes, q.elem);
  return q;
}

// tractor::Program  ->  trxfile::Program   (program.h:39-303)
inline trxfile::Program convertProgram(const tractor::Program &src) {
  trxfile::Program dst;
  dst.memory_size = src.memorySize();
  std::map<const tractor::Operator *, uint32_t> op_index;
  uint32_t lane_width = 1;
  for (auto &inst : src.instructions()) {
    const tractor::Operator *op = inst.op();
    auto it = op_index.find(op);
    if (it == op_index.end()) {
      trxfile::OpDesc d;
      d.name = op->name();
      for (auto &arg : op->arguments()) {
        trxfile::ArgDesc a;
        a.size = (uint32_t)arg.size();
        a.is_input = arg.isInput() ? 1 : 0;
        classify(arg.typeInfo(), a.lanes, a.elem);
        lane_width = std::max<uint32_t>(lane_width, a.lanes);
        d.args.push_back(a);
      }
      it = op_index.emplace(op, (uint32_t)dst.ops.size()).first;
      dst.ops.push_back(d);
    }
    dst.code.push_back(it->second);
    for (auto a : inst.args()) dst.code.push_back((uint64_t)a);
    dst.n_instructions++;
  }
  for (auto &p : src.inputs()) dst.inputs.push_back(convertPort(p));
  for (auto &p : src.parameters()) dst.parameters.push_back(convertPort(p));
  for (auto &p : src.outputs()) dst.outputs.push_back(convertPort(p));
  for (auto &p : src.constants()) dst.constants.push_back(convertPort(p));
  for (auto &v : {&dst.inputs, &dst.parameters, &dst.outputs, &dst.constants})
    for (auto &p : *v) lane_width = std::max<uint32_t>(lane_width, p.lanes);
  for (auto &g : src.goals()) dst.goals.push_back({(uint64_t)g.port(), (int32_t)g.priority()});
  dst.const_data.assign(src.constData().begin(), src.constData().end());
  dst.lane_width = lane_width;
  return dst;
}

// The five programs a solver derives from one recording
// (tractor/include/tractor/solvers/base.h:112-138, gradients.cpp:22-407).
struct GradientSet {
  tractor::Program prog, prep, fprop, bprop, hprop, accu;
  void build() { tractor::buildGradients(prog, prep, &fprop, &bprop, &hprop, &accu); }
  void addTo(trxfile::Bundle &b, const std::string &prefix = "") const {
    b.programs.emplace_back(prefix + "prog", convertProgram(prog));
    b.programs.emplace_back(prefix + "prep", convertProgram(prep));
    b.programs.emplace_back(prefix + "fprop", convertProgram(fprop));
    b.programs.emplace_back(prefix + "bprop", convertProgram(bprop));
    b.programs.emplace_back(prefix + "hprop", convertProgram(hprop));
    b.programs.emplace_back(prefix + "accu", convertProgram(accu));
  }
};

template <class Ports> size_t portElements(const Ports &ports) {
  size_t n = 0;
  for (auto &p : ports) n += p.size() / sizeof(double);
  return n;
}

// One Memory + the executables, driven the way solvers/gd.h:43-63 drives them.
struct RefRunner {
  std::shared_ptr<tractor::Engine> engine;
  std::shared_ptr<tractor::Executable> x_prog, x_prep, x_fprop, x_bprop, x_hprop, x_accu;
  std::shared_ptr<tractor::Memory> memory;
  const GradientSet *gs = nullptr;

  explicit RefRunner(const GradientSet &g, bool loop_engine = false) : gs(&g) {
    if (loop_engine) engine = std::make_shared<tractor::LoopEngine>();
    else engine = std::make_shared<tractor::SimpleEngine>();
    x_prog = engine->compile(g.prog);
    x_prep = engine->compile(g.prep);
    x_fprop = engine->compile(g.fprop);
    x_bprop = engine->compile(g.bprop);
    x_hprop = engine->compile(g.hprop);
    x_accu = engine->compile(g.accu);
    memory = engine->createMemory();
  }
  void parameterize(const std::vector<double> &p) {
    if (!gs->prog.parameters().empty()) x_prog->parameterVector(p, memory);
  }
  // r = f(x)
  void forward(const std::vector<double> &x, std::vector<double> &r) { x_prog->run(x, memory, r); }
  void prepare() { x_prep->execute(memory); }
  // g = J^T v
  void reverse(const std::vector<double> &v, std::vector<double> &g) { x_bprop->run(v, memory, g); }
  // dr = J dx
  void tangent(const std::vector<double> &dx, std::vector<double> &dr) { x_fprop->run(dx, memory, dr); }
  // h = J^T J dx
  void gaussNewton(const std::vector<double> &dx, std::vector<double> &h) { x_hprop->run(dx, memory, h); }
};

struct MuteCout {
  std::streambuf *old;
  std::ostringstream sink;
  MuteCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~MuteCout() { std::cout.rdbuf(old); }
};

}  // namespace refc
