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
#include <string.h>

#include "trx_ops.h"
#include "vm_math.cuh"
#include "collision/trx_collide.h"

namespace trxvm {

// Argument words.  bit 31 set: the value lives in this warp's on-chip slot ring (shared memory),
// bits [0,16) = offset in doubles; otherwise the word is the slot index (byte address / 8) in the
// tile's memory image.  "dual" records carry a second row of words per argument: the memory-image
// index an output is written to in addition to its ring slot (kNoSlot = none).
static constexpr uint32_t kRingFlag = 0x80000000u;
static constexpr uint32_t kNoSlot = 0xffffffffu;

// kAS / kCS: argument stride and component stride as compile-time constants (0 = run-time fields `as`
// and `cs`).  kPlain: every argument word is a memory-image slot (no ring, no pool, no dual rows), so
// an operand address is  m + 8 * (word + lane)  plus an immediate per component.
template <int kAS, int kCS, bool kPlain>
struct CtxT {
  static constexpr bool kPlainWords = kPlain;
  double *m;           // base of this tile's memory image (global memory)
  double *ring;        // base of this warp's slot ring (shared memory)
  const uint32_t *a;   // first argument word of this thread's sub-instruction
  int as;              // stride between consecutive arguments in the record (4 or 32)
  int cs;              // component stride in doubles (8 lane-typed, 1 scalar)
  int lo;              // lane offset (0..7)
  int nargs;           // number of arguments (rows) of the record
  bool dual;           // record has a second row of words per argument
  bool rings;          // false: every argument word is a memory-image slot (generic executables)
  const double *gpool; // non-null: words with bit 31 set index the program's deduplicated constant pool
                       // in global memory (small, shared by all tiles, L1-resident) instead of the image
  bool preloaded;      // argument words were fetched into w[] when the record started
  uint32_t w[12];      // (registers: every use indexes it with a compile-time constant)
  uint32_t w2[12];     // second row of a dual record: memory-image slot of an output that also lives on chip
  TRX_HD int AS() const { return kAS ? kAS : as; }
  TRX_HD int CS() const { return kCS ? kCS : cs; }
  TRX_HD int LO() const { return kCS == 1 ? 0 : lo; }
  TRX_HD bool RINGS() const { return kPlain ? false : rings; }
  TRX_HD bool DUAL() const { return kPlain ? false : dual; }
  TRX_HD const double *GPOOL() const { return kPlain ? nullptr : gpool; }
};
using Ctx = CtxT<0, 0, false>;
using CtxPlain = CtxT<0, 0, true>;  // run-time strides, plain argument words (host emulation of the staged kernel)

template <class C>
TRX_HD uint32_t argword(const C &c, int k) { return c.preloaded ? c.w[k] : c.a[k * c.AS()]; }
// all argument words of the record at once: independent loads, one latency instead of one per operand
template <class C>
TRX_HD void preload_args(C &c) {
#pragma unroll
  for (int k = 0; k < 12; k++)
    if (k < c.nargs) c.w[k] = c.a[k * c.AS()];
  if (c.RINGS() && c.DUAL()) {
#pragma unroll
    for (int k = 0; k < 12; k++)
      if (k < c.nargs) c.w2[k] = c.a[(c.nargs + k) * c.AS()];
  }
  c.preloaded = true;
}
// the same for an operator whose argument count is known where it is expanded (exec below)
template <int kN, class C>
TRX_HD void preload_n(C &c) {
  if (c.preloaded) return;
#pragma unroll
  for (int k = 0; k < kN; k++) c.w[k] = c.a[k * c.AS()];
  if (c.RINGS() && c.DUAL()) {
#pragma unroll
    for (int k = 0; k < kN; k++) c.w2[k] = c.a[(kN + k) * c.AS()];
  }
  c.preloaded = true;
}
template <class C>
TRX_HD double *slotptr(const C &c, uint32_t w) {
  if (!c.RINGS()) {
    // gpool is a compile-time null in kernels instantiated without a pool: the test folds away
    if (c.GPOOL() != nullptr && (w & kRingFlag)) return const_cast<double *>(c.GPOOL()) + (w & 0x7fffffffu);
    return c.m + w;
  }
  return (w & kRingFlag) ? (c.ring + (w & 0xffffu)) : (c.m + w);
}
template <class C>
TRX_HD double ld(const C &c, int k, int j) {
  // plain words: slot index + lane in 32 bits (a tile image is < 32 GB), one widening multiply-add to the
  // address, the component offset as an immediate
  if (C::kPlainWords) return (c.m + (uint32_t)(argword(c, k) + (uint32_t)c.LO()))[j * c.CS()];
  return slotptr(c, argword(c, k))[j * c.CS() + c.LO()];
}
template <class C>
TRX_HD void st(const C &c, int k, int j, double v) {
  if (C::kPlainWords) {
    (c.m + (uint32_t)(argword(c, k) + (uint32_t)c.LO()))[j * c.CS()] = v;
    return;
  }
  slotptr(c, argword(c, k))[j * c.CS() + c.LO()] = v;
  if (c.RINGS() && c.DUAL()) {
    uint32_t w2 = c.preloaded ? c.w2[k] : c.a[(c.nargs + k) * c.AS()];
    if (w2 != kNoSlot) c.m[w2 + j * c.CS() + c.LO()] = v;
  }
}
// BatchScalar-typed argument: one double shared by the lanes (always in the memory image)
template <class C>
TRX_HD double lds(const C &c, int k) { return c.m[argword(c, k)]; }
// reverse of `batch`: da = sum of the lanes of dx, left to right (dexopt/src/ops.h:22-30)
template <class C>
TRX_HD void sts_lanesum(const C &c, int kout, int kin) {
  if (c.LO() != 0) return;
  const double *p = slotptr(c, argword(c, kin));
  double rs = 0.0;
  int n = (c.CS() == 8) ? 8 : 1;
  for (int i = 0; i < n; i++) rs += p[i];
  c.m[argword(c, kout)] = rs;
}
template <class C>
TRX_HD V3 ldv(const C &c, int k, int o) { return V3{ld(c, k, o), ld(c, k, o + 1), ld(c, k, o + 2)}; }
template <class C>
TRX_HD Q4 ldq(const C &c, int k, int o) { return Q4{ld(c, k, o), ld(c, k, o + 1), ld(c, k, o + 2), ld(c, k, o + 3)}; }
template <class C>
TRX_HD PoseV ldp(const C &c, int k, int o) {
  PoseV p;
  p.t = ldv(c, k, o);
  p.r = ldq(c, k, o + 3);
  return p;
}
template <class C>
TRX_HD TwistV ldt(const C &c, int k, int o) {
  TwistV t;
  t.t = ldv(c, k, o);
  t.r = ldv(c, k, o + 3);
  return t;
}
template <class C>
TRX_HD void stv(const C &c, int k, int o, const V3 &v) {
  st(c, k, o, v.x);
  st(c, k, o + 1, v.y);
  st(c, k, o + 2, v.z);
}
template <class C>
TRX_HD void stq(const C &c, int k, int o, const Q4 &q) {
  st(c, k, o, q.x);
  st(c, k, o + 1, q.y);
  st(c, k, o + 2, q.z);
  st(c, k, o + 3, q.w);
}
template <class C>
TRX_HD void stp(const C &c, int k, int o, const PoseV &p) {
  stv(c, k, o, p.t);
  stq(c, k, o + 3, p.r);
}
template <class C>
TRX_HD void stt(const C &c, int k, int o, const TwistV &t) {
  stv(c, k, o, t.t);
  stv(c, k, o + 3, t.r);
}

// BatchScalar-typed output (one double for all lanes): the first lane stores it
template <class C>
TRX_HD void sts(const C &c, int k, double v) {
  if (c.LO() == 0) c.m[argword(c, k)] = v;
}

// ---- in-tape random numbers (dexopt/src/ops.h:120-204: add_random_normal, add_random_uniform, dropout2)
// The reference draws from a thread_local std::mt19937 seeded from rand() -- fresh, unreproducible values at every
// execute.  Here a draw is a pure function of (seed, execute counter, rollout, instruction): Philox4x32-10
// (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC11) with
//   counter = {rollout lo, rollout hi, output slot of the instruction, draw index}, key = {seed ^ epoch}.
// The host sets the state before every execute of a tape that draws (trx_engine.cu executeImpl).
struct RngState {
  unsigned long long seed = 0x9E3779B97F4A7C15ull, epoch = 0;
  const double *base = nullptr;        // first tile of the memory the program runs on
  unsigned long long tile_stride = 1;  // doubles
  unsigned long long lane_base = 0;    // rollout of lane 0 of tile 0
};
#ifdef __CUDACC__
__device__ RngState g_rng_dev;
#endif
// host executions of the operators (tests/emu): one state per process, whichever copy of these inline functions
// the dynamic linker picks
inline RngState &rngHostState() {
  static RngState state;
  return state;
}

TRX_HD void philox4x32_10(const uint32_t (&ctr)[4], const uint32_t (&key)[2], uint32_t (&out)[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; round++) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1,
                   n3 = (uint32_t)p0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
// two uniforms in [0, 1) with 53 random bits each, for output argument kout of the running instruction
template <class C>
TRX_HD void rng_uniform2(const C &c, int kout, double &u0, double &u1) {
#ifdef __CUDA_ARCH__
  const RngState &r = g_rng_dev;
#else
  const RngState &r = rngHostState();
#endif
  const unsigned long long tile = (unsigned long long)(c.m - r.base) / r.tile_stride;
  const unsigned long long rollout = r.lane_base + tile * 8 + (unsigned)c.LO();
  const unsigned long long k = r.seed ^ (r.epoch * 0xD1B54A32D192ED03ull);
  const uint32_t ctr[4] = {(uint32_t)rollout, (uint32_t)(rollout >> 32), argword(c, kout), 0u};
  const uint32_t key[2] = {(uint32_t)k, (uint32_t)(k >> 32)};
  uint32_t o[4];
  philox4x32_10(ctr, key, o);
  u0 = (double)((((unsigned long long)o[0] << 32) | o[1]) >> 11) * (1.0 / 9007199254740992.0);
  u1 = (double)((((unsigned long long)o[2] << 32) | o[3]) >> 11) * (1.0 / 9007199254740992.0);
}
template <class C>
TRX_HD double rng_uniform(const C &c, int kout) {
  double u0, u1;
  rng_uniform2(c, kout, u0, u1);
  return u0;
}
// standard normal (Box-Muller on one Philox block)
template <class C>
TRX_HD double rng_normal(const C &c, int kout) {
  double u0, u1;
  rng_uniform2(c, kout, u0, u1);
  return sqrt(-2.0 * log(1.0 - u0)) * cos(6.283185307179586476925286766559 * u1);
}

// ---- collision shapes (collision/trx_collide.h): the table the host registered, mirrored to the device before a
// program with collision operators runs (trx_engine.cu executeImpl).  uint64-typed arguments are one slot per
// instruction holding the integer's bits.
#ifdef __CUDACC__
__device__ const double *g_shape_heap_dev;
#endif
inline const double *&shapeHeapHost() {
  static const double *heap = nullptr;
  return heap;
}
TRX_HD const double *shapeHeap() {
#ifdef __CUDA_ARCH__
  return g_shape_heap_dev;
#else
  return shapeHeapHost();
#endif
}
template <class C>
TRX_HD uint64_t ldu(const C &c, int k) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(c.m[argword(c, k)]);
#else
  uint64_t bits;
  memcpy(&bits, &c.m[argword(c, k)], sizeof(bits));
  return bits;
#endif
}

// number of arguments in a signature string (tokens separated by single blanks)
TRX_HD constexpr int sigArgCount(const char *sig) {
  int n = 0;
  bool in_token = false;
  for (const char *p = sig; *p; p++) {
    if (*p == ' ') in_token = false;
    else if (!in_token) {
      in_token = true;
      n++;
    }
  }
  return n;
}

// One instruction.  `opcode` is warp-uniform, so the switch does not diverge.  The argument words
// are fetched first, as many as the operator's signature names.
// kWithCollision = false compiles the collision operators' bodies out: their GJK / EPA routines are real calls with
// a 7 KB stack frame, which costs every launch of a kernel that contains them local memory and registers around the
// call sites -- programs without those operators run an instantiation without them (trx_engine.cu executeImpl).
template <bool kWithCollision = true, class C>
TRX_HD void exec(uint32_t opcode, C &c) {
  switch (opcode) {
#define TRX_OP(ENUM, NAME, MODE, SIG, ...)        \
  case trx::OP_##ENUM: {                          \
    constexpr int kNumArgs = sigArgCount(SIG);    \
    preload_n<kNumArgs>(c);                       \
    __VA_ARGS__                                   \
  } break;
#include "vm_ops.def"
#undef TRX_OP
    default:
      break;
  }
}


// Block-cooperative operators.  `a` = argument words (slot indices in doubles) followed by the two
// immediates n_in, n_out; thread `tid` of `nthreads` takes a strided share of the work.
TRX_HD void exec_coop(uint32_t opcode, const uint32_t *a, double *m, int tid, int nthreads) {
  switch (opcode) {
    case trx::OP_X_DENSE: {
      const double *x = m + a[0], *W = m + a[1], *b = m + a[2];
      double *y = m + a[3];
      const int n_in = (int)a[4], n_out = (int)a[5];
      for (int o = tid; o < n_out * 8; o += nthreads) {
        const int j = o >> 3, lane = o & 7;
        double acc = 0.0;
        int row = 0;
        // tensor.h:147-159: rows in groups of four through dot4add, then madd for the tail
        for (; row + 3 < n_in; row += 4) {
          acc = x[(row + 0) * 8 + lane] * W[(row + 0) * n_out + j] + x[(row + 1) * 8 + lane] * W[(row + 1) * n_out + j] +
                x[(row + 2) * 8 + lane] * W[(row + 2) * n_out + j] + x[(row + 3) * 8 + lane] * W[(row + 3) * n_out + j] + acc;
        }
        for (; row < n_in; row++) acc = x[row * 8 + lane] * W[row * n_out + j] + acc;
        y[j * 8 + lane] = acc + b[j];  // neural.h:176-181
      }
    } break;
    case trx::OP_R_X_DENSE: {
      const double *x = m + a[0], *W = m + a[1];
      double *gx = m + a[4], *gW = m + a[5], *gb = m + a[6];
      const double *gy = m + a[7];
      const int n_in = (int)a[8], n_out = (int)a[9];
      // gx[i] = sum_j gy[j] * W[i][j]
      for (int o = tid; o < n_in * 8; o += nthreads) {
        const int i = o >> 3, lane = o & 7;
        double acc = 0.0;
        for (int j = 0; j < n_out; j++) acc += gy[j * 8 + lane] * W[i * n_out + j];
        gx[i * 8 + lane] = acc;
      }
      // gW[i][j] += sum over lanes (left to right, as reverse_batch) of x[i] * gy[j]
      for (int e = tid; e < n_in * n_out; e += nthreads) {
        const int i = e / n_out, j = e - i * n_out;
        double rs = 0.0;
        for (int lane = 0; lane < 8; lane++) rs += gy[j * 8 + lane] * x[i * 8 + lane];
        gW[e] += rs;
      }
      for (int j = tid; j < n_out; j += nthreads) {
        double rs = 0.0;
        for (int lane = 0; lane < 8; lane++) rs += gy[j * 8 + lane];
        gb[j] += rs;
      }
    } break;
    case trx::OP_F_X_DENSE: {
      // tangent of a dense layer: dy = dx W + x dW + db
      const double *x = m + a[0], *W = m + a[1];
      const double *dx = m + a[4], *dW = m + a[5], *db = m + a[6];
      double *dy = m + a[7];
      const int n_in = (int)a[8], n_out = (int)a[9];
      for (int o = tid; o < n_out * 8; o += nthreads) {
        const int j = o >> 3, lane = o & 7;
        double acc = 0.0;
        for (int i = 0; i < n_in; i++) acc += dx[i * 8 + lane] * W[i * n_out + j] + x[i * 8 + lane] * dW[i * n_out + j];
        dy[j * 8 + lane] = acc + db[j];
      }
    } break;
    case trx::OP_X_ZERO_BLOCK: {
      double *z = m + a[0];
      const int n = (int)a[1];
      for (int e = tid; e < n; e += nthreads) z[e] = 0.0;
    } break;
    default:
      break;
  }
}

}  // namespace trxvm
