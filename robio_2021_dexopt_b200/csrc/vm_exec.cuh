// vm_exec.cuh -- the tape VM's instruction executor (device side).
//
// Replaces the reference's per-instruction indirect call
//     for (auto &inst : _instructions) inst.op(temp.data(), _arguments.data() + inst.base);
// (tractor/src/engines/simple.cpp:32-34) by a warp-uniform opcode switch whose
// bodies operate on structure-of-arrays lane tiles:
//   slot component j of argument k, lane l  ->  m[a_k + j * cs + l]
// with cs = 8 for lane-typed (*_8d) operators and cs = 1 for scalar (*_d) ones.
#pragma once
#include <stdint.h>

#include "trx_ops.h"
#include "vm_math.cuh"

namespace trxvm {

struct Ctx {
  double *m;           // base of this tile's memory image
  const uint32_t *a;   // first argument word of this thread's sub-instruction
  int as;              // stride between consecutive arguments in the record (4 or 32)
  int cs;              // component stride in doubles (8 lane-typed, 1 scalar)
  int lo;              // lane offset (0..7)
};

TRX_HD uint32_t argbase(const Ctx &c, int k) { return c.a[k * c.as]; }
TRX_HD double ld(const Ctx &c, int k, int j) { return c.m[argbase(c, k) + j * c.cs + c.lo]; }
TRX_HD void st(const Ctx &c, int k, int j, double v) { c.m[argbase(c, k) + j * c.cs + c.lo] = v; }
// BatchScalar-typed argument: one double shared by the lanes
TRX_HD double lds(const Ctx &c, int k) { return c.m[argbase(c, k)]; }
// reverse of `batch`: da = sum of the lanes of dx, left to right (dexopt/src/ops.h:22-30)
TRX_HD void sts_lanesum(const Ctx &c, int kout, int kin) {
  if (c.lo != 0) return;
  const double *p = c.m + argbase(c, kin);
  double rs = 0.0;
  int n = (c.cs == 8) ? 8 : 1;
  for (int i = 0; i < n; i++) rs += p[i];
  c.m[argbase(c, kout)] = rs;
}
TRX_HD V3 ldv(const Ctx &c, int k, int o) { return V3{ld(c, k, o), ld(c, k, o + 1), ld(c, k, o + 2)}; }
TRX_HD Q4 ldq(const Ctx &c, int k, int o) { return Q4{ld(c, k, o), ld(c, k, o + 1), ld(c, k, o + 2), ld(c, k, o + 3)}; }
TRX_HD PoseV ldp(const Ctx &c, int k, int o) {
  PoseV p;
  p.t = ldv(c, k, o);
  p.r = ldq(c, k, o + 3);
  return p;
}
TRX_HD TwistV ldt(const Ctx &c, int k, int o) {
  TwistV t;
  t.t = ldv(c, k, o);
  t.r = ldv(c, k, o + 3);
  return t;
}
TRX_HD void stv(const Ctx &c, int k, int o, const V3 &v) {
  st(c, k, o, v.x);
  st(c, k, o + 1, v.y);
  st(c, k, o + 2, v.z);
}
TRX_HD void stq(const Ctx &c, int k, int o, const Q4 &q) {
  st(c, k, o, q.x);
  st(c, k, o + 1, q.y);
  st(c, k, o + 2, q.z);
  st(c, k, o + 3, q.w);
}
TRX_HD void stp(const Ctx &c, int k, int o, const PoseV &p) {
  stv(c, k, o, p.t);
  stq(c, k, o + 3, p.r);
}
TRX_HD void stt(const Ctx &c, int k, int o, const TwistV &t) {
  stv(c, k, o, t.t);
  stv(c, k, o + 3, t.r);
}

// One instruction.  `opcode` is warp-uniform, so the switch does not diverge.
TRX_HD void exec(uint32_t opcode, const Ctx &c) {
  switch (opcode) {
#define TRX_OP(ENUM, NAME, MODE, SIG, ...) \
  case trx::OP_##ENUM: {                   \
    __VA_ARGS__                            \
  } break;
#include "vm_ops.def"
#undef TRX_OP
    default:
      break;
  }
}

}  // namespace trxvm
