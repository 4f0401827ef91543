// trx_engine.cu -- the B200 tape execution engine behind include/trx_b200.h.
//
// One CTA executes one lane tile (8 rollouts = one reference memory image,
// tractor/include/tractor/engines/base.h:13-31) of the tape; the W warps of the
// CTA walk W pre-scheduled instruction streams (trx_lower.h) and meet at a
// block barrier after every dependency level.  Tiles never interact: rollouts
// are independent in the reference too (the only cross-lane operator is the
// reverse of `batch`, dexopt/src/ops.h:22-34, which stays inside a tile here;
// the sum over tiles happens when a reverse program's scalar outputs are read).
//
// Compiled for sm_100a only.  No CPU path: every entry point that needs the
// device reports TRX_ERR_CUDA when there is none.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <limits>
#include <cmath>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/trx_b200.h"
#include "trx_lower.h"
#include "trx_combine.h"
#ifdef TRX_WITH_EXPERIMENTS
#include "trx_fuse.h"  // chain scheduling with on-chip slot rings: measured slower, not in the product library
#endif
#include "trx_direct.h"
#include "vm_exec.cuh"
#include "trx_comm.h"
#include "collision/trx_shape_table.h"
#include "trx_rollout.h"

namespace {

long long *g_trace = nullptr;  // development aid (trx_debug_trace)
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

// development knob: TRX_STAGED=0 runs generic executables on the kernel that reads its stream from global memory
constexpr int kStagedWarps = 20;  // warps per tile of the staged kernel
bool staged_streams() {
  static int v = getenv("TRX_STAGED") ? atoi(getenv("TRX_STAGED")) : 1;
  return v != 0;
}

trx_status fail(trx_status code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

#define TRX_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(TRX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
  } while (0)

// ------------------------------------------------------------------ kernels


// Block-cooperative dense layer on the device: operand blocks staged in shared memory, several
// independent accumulators per thread (the host/emulation statement of the same operators is
// trxvm::exec_coop in vm_exec.cuh).  Called by every thread of the CTA at the same stream position.
__device__ __forceinline__ void dense_forward_dev(const uint32_t *a, double *m, double *sx) {
  const double *x = m + a[0], *W = m + a[1], *b = m + a[2];
  double *y = m + a[3];
  const int n_in = (int)a[4], n_out = (int)a[5];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < n_in * 8; e += nt) sx[e] = x[e];
  __syncthreads();
  for (int o = tid; o < n_out * 8; o += nt) {
    const int j = o >> 3, lane = o & 7;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int row = 0;
#pragma unroll 2
    for (; row + 3 < n_in; row += 4) {
      acc0 = fma(sx[(row + 0) * 8 + lane], __ldg(W + (row + 0) * n_out + j), acc0);
      acc1 = fma(sx[(row + 1) * 8 + lane], __ldg(W + (row + 1) * n_out + j), acc1);
      acc2 = fma(sx[(row + 2) * 8 + lane], __ldg(W + (row + 2) * n_out + j), acc2);
      acc3 = fma(sx[(row + 3) * 8 + lane], __ldg(W + (row + 3) * n_out + j), acc3);
    }
    for (; row < n_in; row++) acc0 = fma(sx[row * 8 + lane], __ldg(W + row * n_out + j), acc0);
    y[o] = ((acc0 + acc1) + (acc2 + acc3)) + __ldg(b + j);
  }
  __syncthreads();  // sx is reused by the next cooperative record
}

__device__ __forceinline__ void dense_reverse_dev(const uint32_t *a, double *m, double *sx, double *sg, double *sgT) {
  const double *x = m + a[0], *W = m + a[1];
  double *gx = m + a[4], *gW = m + a[5], *gb = m + a[6];
  const double *gy = m + a[7];
  const int n_in = (int)a[8], n_out = (int)a[9];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < n_in * 8; e += nt) sx[e] = x[e];
  // gy twice: [j][lane] for the gx products (a warp reads the 8 lanes of one j), [lane][j] for the lane sums
  // of gW and gb (a warp reads consecutive j of one lane) -- either use on the other layout is a 16-way
  // shared-memory bank conflict
  for (int e = tid; e < n_out * 8; e += nt) {
    const double v = gy[e];
    sg[e] = v;
    sgT[(e & 7) * n_out + (e >> 3)] = v;
  }
  __syncthreads();
  // gx[i] = sum_j gy[j] * W[i][j]
  for (int o = tid; o < n_in * 8; o += nt) {
    const int i = o >> 3, lane = o & 7;
    const double *Wi = W + i * n_out;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int j = 0;
#pragma unroll 2
    for (; j + 3 < n_out; j += 4) {
      acc0 = fma(sg[(j + 0) * 8 + lane], __ldg(Wi + j + 0), acc0);
      acc1 = fma(sg[(j + 1) * 8 + lane], __ldg(Wi + j + 1), acc1);
      acc2 = fma(sg[(j + 2) * 8 + lane], __ldg(Wi + j + 2), acc2);
      acc3 = fma(sg[(j + 3) * 8 + lane], __ldg(Wi + j + 3), acc3);
    }
    for (; j < n_out; j++) acc0 = fma(sg[j * 8 + lane], __ldg(Wi + j), acc0);
    gx[o] = (acc0 + acc1) + (acc2 + acc3);
  }
  // gW[i][j] += sum over lanes of x[i] * gy[j]   (lane sum as reverse_batch, then one accumulate per frame).
  // Four elements per thread and trip: their read-modify-writes are independent, so the four loads of
  // gW are issued together instead of one L2 round trip per element.
  auto lanedot = [&](int e) {
    const int i = e / n_out, j = e - i * n_out;
    const double *xi = sx + i * 8, *gj = sgT + j;
    double r0 = xi[0] * gj[0], r1 = xi[1] * gj[n_out];
    r0 = fma(xi[2], gj[2 * n_out], r0);
    r1 = fma(xi[3], gj[3 * n_out], r1);
    r0 = fma(xi[4], gj[4 * n_out], r0);
    r1 = fma(xi[5], gj[5 * n_out], r1);
    r0 = fma(xi[6], gj[6 * n_out], r0);
    r1 = fma(xi[7], gj[7 * n_out], r1);
    return r0 + r1;
  };
  const int n_w = n_in * n_out;
  int e = tid;
  for (; e + 3 * nt < n_w; e += 4 * nt) {
    const double w0 = gW[e], w1 = gW[e + nt], w2 = gW[e + 2 * nt], w3 = gW[e + 3 * nt];
    const double d0 = lanedot(e), d1 = lanedot(e + nt), d2 = lanedot(e + 2 * nt), d3 = lanedot(e + 3 * nt);
    gW[e] = w0 + d0;
    gW[e + nt] = w1 + d1;
    gW[e + 2 * nt] = w2 + d2;
    gW[e + 3 * nt] = w3 + d3;
  }
  for (; e < n_w; e += nt) gW[e] += lanedot(e);
  for (int j = tid; j < n_out; j += nt) {
    const double *gj = sgT + j;
    gb[j] += ((gj[0] + gj[n_out]) + (gj[2 * n_out] + gj[3 * n_out])) +
             ((gj[4 * n_out] + gj[5 * n_out]) + (gj[6 * n_out] + gj[7 * n_out]));
  }
  __syncthreads();
}

// ---- dense layers on the FP64 tensor cores (DMMA m8n8k4; tcgen05 has no f64 kind) ----------------
// The 8 rollouts of a tile are the M = 8 rows of every product, so a dense layer is a row of 8x8 output
// tiles, one warp each.  Fragment layout of mma.sync.m8n8k4.f64 (g = lane >> 2, t = lane & 3):
//   A (8x4, row)  a    = A[g][t]          B (4x8, col)  b = B[t][g]          C (8x8)  c0,c1 = C[g][2t], C[g][2t+1]
// Each thread loads 2 operands per 8 multiply-adds instead of 2 per 1.
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One 8x8 output tile: c += A * B over k = 0 .. k_pad (multiple of 4), eight k-steps (32 values of k) per
// trip so that their operand loads are in flight together, on two accumulator chains.  loadA(k) / loadB(k)
// return the fragment elements for this thread's k (A from shared memory, B from global memory, zero
// when out of range).  Kept small on purpose: these functions are inlined into the interpreter and must
// not push it over its register budget (a 124-register variant with operands prefetched ahead of the
// staging barrier was measured: no faster, and the spilling one 12 % slower overall).
template <class FA, class FB>
__device__ __forceinline__ void mma_tile(double &c0, double &c1, int k_pad, int t, FA loadA, FB loadB) {
  double d0 = 0.0, d1 = 0.0;
  for (int k0 = 0; k0 < k_pad; k0 += 32) {
    double fa[8], fb[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const bool on = k0 + 4 * u < k_pad;
      fb[u] = on ? loadB(k0 + 4 * u + t) : 0.0;
      fa[u] = on ? loadA(k0 + 4 * u + t) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
      if (k0 + 4 * u < k_pad) dmma884(c0, c1, fa[u], fb[u]);
      if (k0 + 4 * u + 4 < k_pad) dmma884(d0, d1, fa[u + 1], fb[u + 1]);
    }
  }
  c0 += d0;
  c1 += d1;
}

// y[j][lane] = sum_i x[i][lane] W[i][j] + b[j]      (neural.h:176-181, tensor.h:147-159)
__device__ __forceinline__ void dense_forward_mma(const uint32_t *a, double *m, double *sx) {
  const double *x = m + a[0], *W = m + a[1], *b = m + a[2];
  double *y = m + a[3];
  const int n_in = (int)a[4], n_out = (int)a[5];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int k_pad = (n_in + 3) & ~3;
  for (int e = tid; e < k_pad * 8; e += nt) sx[e] = (e < n_in * 8) ? x[e] : 0.0;
  __syncthreads();
  const int lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int tile = tid >> 5; tile * 8 < n_out; tile += nt >> 5) {
    const int n = tile * 8 + g;  // output column this thread feeds as B operand
    const double *Wn = W + n;
    double c0 = 0.0, c1 = 0.0;
    mma_tile(
        c0, c1, k_pad, t, [&](int k) { return sx[k * 8 + g]; },
        [&](int k) { return (k < n_in && n < n_out) ? __ldg(Wn + (size_t)k * n_out) : 0.0; });
    const int j = tile * 8 + 2 * t;
    if (j < n_out) y[j * 8 + g] = c0 + __ldg(b + j);
    if (j + 1 < n_out) y[(j + 1) * 8 + g] = c1 + __ldg(b + j + 1);
  }
  __syncthreads();  // sx is reused by the next cooperative record
}

// gx[i][lane] = sum_j gy[j][lane] W[i][j];  gW[i][j] += sum_lanes x[i] gy[j];  gb[j] += sum_lanes gy[j]
// clk (development aid, tracing instantiation only): per-phase clock sums of tile 0
__device__ __forceinline__ void dense_reverse_mma(const uint32_t *a, double *m, double *sx, double *sg, long long *clk = nullptr) {
  long long tc[6];
  if (clk) tc[0] = tc[1] = clock64();
  const double *x = m + a[0], *W = m + a[1];
  double *gx = m + a[4], *gW = m + a[5], *gb = m + a[6];
  const double *gy = m + a[7];
  const int n_in = (int)a[8], n_out = (int)a[9];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int in_pad = (n_in + 7) & ~7, out_pad = (n_out + 7) & ~7;
  for (int e = tid; e < in_pad * 8; e += nt) sx[e] = (e < n_in * 8) ? x[e] : 0.0;
  for (int e = tid; e < out_pad * 8; e += nt) sg[e] = (e < n_out * 8) ? gy[e] : 0.0;
  __syncthreads();
  if (clk) tc[2] = clock64();
  const int lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int warp = tid >> 5, n_warps = nt >> 5;
  const int tiles_in = in_pad >> 3, tiles_out = out_pad >> 3;
  // gx: rows = rollouts, columns = i, k = j.   A[g][t] = gy[k0+t][g],  B[t][g] = W[i0+g][k0+t]
  for (int tile = warp; tile < tiles_in; tile += n_warps) {
    const int i = tile * 8 + g;
    const double *Wi = W + (size_t)i * n_out;
    double c0 = 0.0, c1 = 0.0;
    mma_tile(
        c0, c1, out_pad, t, [&](int k) { return sg[k * 8 + g]; },
        [&](int k) { return (k < n_out && i < n_in) ? __ldg(Wi + k) : 0.0; });
    const int i0 = tile * 8 + 2 * t;
    if (i0 < n_in) gx[i0 * 8 + g] = c0;
    if (i0 + 1 < n_in) gx[(i0 + 1) * 8 + g] = c1;
  }
  if (clk) tc[3] = clock64();
  // gW: rows = i, columns = j, k = rollout lane.   A[g][t] = x[i0+g][l0+t],  B[t][g] = gy[j0+g][l0+t]
  // The accumulator starts from what the earlier frames left in gW.  Tiles (ti, tj) are walked without a
  // division per tile; two tiles per trip so that their read-modify-writes overlap.
  {
    int ti = warp / tiles_out, tj = warp - ti * tiles_out;
    const int step_i = n_warps / tiles_out, step_j = n_warps - step_i * tiles_out;
    auto advance = [&](int &vi, int &vj) {
      vi += step_i;
      vj += step_j;
      if (vj >= tiles_out) {
        vj -= tiles_out;
        vi++;
      }
    };
    while (ti < tiles_in) {
      int ui = ti, uj = tj;
      advance(ui, uj);
      const bool second = ui < tiles_in;
      const int i_a = ti * 8 + g, j_a = tj * 8 + 2 * t, i_b = ui * 8 + g, j_b = uj * 8 + 2 * t;
      double *pa = gW + (size_t)i_a * n_out + j_a, *pb = gW + (size_t)i_b * n_out + j_b;
      const bool a0 = i_a < n_in && j_a < n_out, a1 = i_a < n_in && j_a + 1 < n_out;
      const bool b0 = second && i_b < n_in && j_b < n_out, b1 = second && i_b < n_in && j_b + 1 < n_out;
      double ca0 = a0 ? pa[0] : 0.0, ca1 = a1 ? pa[1] : 0.0, cb0 = b0 ? pb[0] : 0.0, cb1 = b1 ? pb[1] : 0.0;
      const double *xa = sx + i_a * 8 + t, *ga = sg + (tj * 8 + g) * 8 + t;
      dmma884(ca0, ca1, xa[0], ga[0]);
      dmma884(ca0, ca1, xa[4], ga[4]);
      if (second) {
        const double *xb = sx + i_b * 8 + t, *gb2 = sg + (uj * 8 + g) * 8 + t;
        dmma884(cb0, cb1, xb[0], gb2[0]);
        dmma884(cb0, cb1, xb[4], gb2[4]);
      }
      if (a0) pa[0] = ca0;
      if (a1) pa[1] = ca1;
      if (b0) pb[0] = cb0;
      if (b1) pb[1] = cb1;
      ti = ui;
      tj = uj;
      if (second) advance(ti, tj);
    }
  }
  if (clk) tc[4] = clock64();
  for (int j = tid; j < n_out; j += nt) {
    const double *gj = sg + j * 8;
    gb[j] += ((gj[0] + gj[1]) + (gj[2] + gj[3])) + ((gj[4] + gj[5]) + (gj[6] + gj[7]));
  }
  __syncthreads();
  if (clk && tid == 0) {
    tc[5] = clock64();
    for (int p = 0; p < 5; p++) clk[p] += tc[p + 1] - tc[p];
    clk[7] += 1;
  }
}

__device__ __forceinline__ bool g_rec_trace_flag(const long long *trace) { return trace[(1ll << 20) - 1] != 0; }

// The interpreter.  grid = tiles, block = W warps; dynamic shared memory = W slot rings.
// record header: opcode[0:11) | wide[11] | n_sub-1[12:17) | n_args[17:21) | coop[21] | dual[22]
// kRings = false: generic executables (every slot in the memory image, plain global loads/stores).
// kPool: argument words with bit 31 set index the program's constant pool in global memory.
template <int kMaxWarps, bool kRings, bool kStreamPrefetch, bool kPool>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
    trx_vm_kernel(const uint32_t *__restrict__ stream, const uint64_t *__restrict__ stream_begin, double *mem,
                  uint64_t tile_stride, uint32_t ring_doubles, long long *trace, const double *__restrict__ pool,
                  uint32_t pool_offset, uint32_t pool_doubles, uint32_t page_words) {
  extern __shared__ double trx_rings[];
  __shared__ double dense_x[128 * 8], dense_g[128 * 8], dense_gT[128 * 8];  // staged operand blocks of the dense operators
  uint32_t trace_level = 0, rec_count = 0;
  double *m = mem + (uint64_t)blockIdx.x * tile_stride;
  const int warp = threadIdx.x >> 5;
  const int lane32 = threadIdx.x & 31;
  const uint32_t *pc = stream + stream_begin[warp];
  uint32_t hdr = __ldg(pc);
  trxvm::Ctx c;
  c.m = m;
  // pool == nullptr: one ring per warp (chain-scheduled streams); otherwise one CTA-wide slot cache
  // (ring + constant pool) shared by all warps of the level-scheduled streams
  c.ring = pool ? trx_rings : trx_rings + (size_t)warp * ring_doubles;
  if (kRings && pool) {
    for (uint32_t e = threadIdx.x; e < pool_doubles; e += blockDim.x) trx_rings[pool_offset + e] = pool[e];
    __syncthreads();
  }
  c.rings = kRings;
  c.gpool = (kRings || !kPool) ? nullptr : pool;
  c.dual = false;
  c.preloaded = false;
  while (true) {
    const uint32_t op = hdr & 0x7ffu;
    if (op == trx::OP_END) break;
    if (op == trx::OP_BARRIER) {
      pc += 1;
      hdr = __ldg(pc);
      if (trace && blockIdx.x == 0) {  // development aid: arrival / release clocks per level
        long long t0 = clock64();
        int arrived = __syncthreads_count(1);  // returns a value: the clock read below cannot issue before release
        long long t1;
        asm volatile("{ .reg .u32 d; add.u32 d, %1, 0; mov.u64 %0, %%clock64; }" : "=l"(t1) : "r"(arrived));
        if (lane32 == 0) {
          trace[((size_t)trace_level * (blockDim.x / 32) + warp) * 2] = t0;
          trace[((size_t)trace_level * (blockDim.x / 32) + warp) * 2 + 1] = t1;
        }
        trace_level++;
        continue;
      }
      __syncthreads();
      continue;
    }
    const bool wide = (hdr >> 11) & 1u;
    const int nsub = (int)((hdr >> 12) & 31u) + 1;
    const int nargs = (int)((hdr >> 17) & 15u);
    if (op == trx::OP_PAGE) {
      const uint32_t at = (uint32_t)(pc - (stream + stream_begin[warp]));
      pc += page_words - (at & (page_words - 1));
      hdr = __ldg(pc);
      continue;
    }
    if (op == trx::OP_PREFETCH) {
      // software prefetch record: one 128-byte line of the tile image per lane, into L2
      const uint32_t *next = pc + 33;
      hdr = __ldg(next);
      if (lane32 < nsub) {
        const double *line = m + (size_t)__ldg(pc + 1 + lane32) * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
      }
      pc = next;
      continue;
    }
    if ((hdr >> 21) & 1u) {
      // block-cooperative record: every warp holds it at the same stream position
      const uint32_t *next = pc + 1 + nargs + 2;
      hdr = __ldg(next);
      const uint32_t n0 = __ldg(pc + 1 + nargs), n1 = __ldg(pc + 2 + nargs);
      if (op == trx::OP_X_DENSE && n0 <= 128)
        dense_forward_dev(pc + 1, m, dense_x);
      else if (op == trx::OP_R_X_DENSE && n0 <= 128 && n1 <= 128)
        dense_reverse_dev(pc + 1, m, dense_x, dense_g, dense_gT);
      else
        trxvm::exec_coop(op, pc + 1, m, (int)threadIdx.x, (int)blockDim.x);
      pc = next;
      continue;
    }
    if (kRings) c.dual = (hdr >> 22) & 1u;
    c.nargs = nargs;
    int sub;
    if (wide) {
      sub = lane32;
      c.as = 32;
      c.cs = 1;
      c.lo = 0;
    } else {
      sub = lane32 >> 3;
      c.as = 4;
      c.cs = 8;
      c.lo = lane32 & 7;
    }
    c.a = pc + 1 + sub;
    long long rt0 = 0, rt1 = 0;
    const bool rec_trace = trace && blockIdx.x == 0 && g_rec_trace_flag(trace);
    if (rec_trace) rt0 = clock64();
    trxvm::preload_args(c);
    if (rec_trace) {
      uint32_t sum = 0;
#pragma unroll
      for (int k = 0; k < 12; k++) sum += (k < nargs) ? c.w[k] : 0;
      asm volatile("{ .reg .u32 d; add.u32 d, %1, 0; mov.u64 %0, %%clock64; }" : "=l"(rt1) : "r"(sum));
    }
    const uint32_t *next = pc + 1 + (c.dual ? 2 * nargs : nargs) * c.as;
    hdr = __ldg(next);  // look ahead: the next header is in flight while this record executes
    // the stream is read sequentially: pull the following lines towards L1 (and further ones into L2)
    // while this record executes
    if (kStreamPrefetch) {
      if (lane32 < 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(next + 32 * lane32));
      else if (lane32 < 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(next + 256 + 64 * (lane32 - 4)));
    }
    if (sub < nsub) trxvm::exec<false>(op, c);  // (collision operators: staged kernel only, see executeImpl)
    __syncwarp();  // chain-scheduled streams: the next record of this warp may read what this one wrote
    if (rec_trace) {
      // development aid: per-record breakdown {start, argument words arrived, done, opcode}; the stores of
      // the record are still in flight at `done`, the loads are not
      __threadfence_block();
      long long rt2 = clock64();
      if (lane32 == 0) {
        long long *rec = trace + (1ll << 20) + ((long long)warp * 4096 + (rec_count & 4095)) * 4;
        rec[0] = rt0;
        rec[1] = rt1;
        rec[2] = rt2;
        rec[3] = (long long)op | ((long long)nsub << 16) | ((long long)trace_level << 32);
      }
      rec_count++;
    }
    pc = next;
  }
}


// The interpreter of generic executables with the instruction stream staged in shared memory.
// Every warp walks its own stream; the stream is cut into pages of kPage words (records never straddle
// a page, trx_lower.h) and kPages of them sit in a per-warp ring that cp.async refills kPages-1 pages
// ahead of the interpreter.  Header and argument words then cost a shared-memory read instead of two
// dependent L2 round trips per record; the only global latency left on a record's critical path is its
// operands.
template <int kMaxWarps, int kPage, int kPages, bool kTrace, bool kCollision>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
    trx_vm_staged_kernel(const uint32_t *__restrict__ stream, const uint64_t *__restrict__ stream_begin, double *mem,
                         uint64_t tile_stride, long long *trace) {
  extern __shared__ __align__(16) uint32_t trx_pages[];  // [warp][kPages][kPage]
  __shared__ double dense_x[128 * 8], dense_g[128 * 8];
  constexpr uint32_t kRing = kPage * kPages;
  uint32_t trace_level = 0, rec_count = 0;
  double *m = mem + (uint64_t)blockIdx.x * tile_stride;
  const int warp = threadIdx.x >> 5;
  const int lane32 = threadIdx.x & 31;
  const uint32_t *base = stream + stream_begin[warp];
  const uint32_t n_pages = (uint32_t)((stream_begin[warp + 1] - stream_begin[warp]) / kPage);
  uint32_t *sbuf = trx_pages + (size_t)warp * kRing;
  // one commit group per page, also for pages past the end of the stream, so that the group count
  // the waits rely on stays fixed
  auto fetch = [&](uint32_t page) {
    if (page < n_pages) {
      const uint32_t *src = base + (size_t)page * kPage;
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(sbuf + (page % kPages) * kPage);
#pragma unroll
      for (int i = 0; i < kPage / 128; i++)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (i * 32 + lane32) * 16),
                     "l"(src + (i * 32 + lane32) * 4)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (uint32_t pg = 0; pg < kPages; pg++) fetch(pg);
  asm volatile("cp.async.wait_group %0;" ::"n"(kPages - 1) : "memory");
  __syncwarp();
  uint32_t pos = 0, cur_page = 0;
  while (true) {
    if ((pos / kPage) != cur_page) {
      // the page just left is free: refill it with the page kPages ahead, then make sure the page
      // being entered has arrived
      __syncwarp();
      fetch(cur_page + kPages);
      cur_page++;
      asm volatile("cp.async.wait_group %0;" ::"n"(kPages - 1) : "memory");
      __syncwarp();
    }
    const bool rec_trace = kTrace && trace && blockIdx.x == 0 && g_rec_trace_flag(trace);
    long long rt0 = 0, rt1 = 0;
    if (rec_trace) rt0 = clock64();
    const uint32_t *pc = sbuf + (pos & (kRing - 1));
    const uint32_t hdr = *pc;
    const uint32_t op = hdr & 0x7ffu;
    const bool wide = (hdr >> 11) & 1u;
    const int nsub = (int)((hdr >> 12) & 31u) + 1;
    const int nargs = (int)((hdr >> 17) & 15u);
    // one test keeps the common case (an operator record) clear of the special words
    if (op >= trx::OP_PAGE || ((hdr >> 21) & 1u)) {
    if (op == trx::OP_END) break;
    if (op == trx::OP_BARRIER) {
      pos += 1;
      if (kTrace && trace && blockIdx.x == 0) {  // development aid: arrival / release clocks per level
        long long t0 = clock64();
        int arrived = __syncthreads_count(1);
        long long t1;
        asm volatile("{ .reg .u32 d; add.u32 d, %1, 0; mov.u64 %0, %%clock64; }" : "=l"(t1) : "r"(arrived));
        if (lane32 == 0) {
          trace[((size_t)trace_level * (blockDim.x / 32) + warp) * 2] = t0;
          trace[((size_t)trace_level * (blockDim.x / 32) + warp) * 2 + 1] = t1;
        }
        trace_level++;
        continue;
      }
      __syncthreads();
      continue;
    }
    if (op == trx::OP_PAGE) {
      pos = (pos | (kPage - 1)) + 1;
      continue;
    }
    if (op == trx::OP_PREFETCH) {
      if (lane32 < nsub) {
        const double *line = m + (size_t)pc[1 + lane32] * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
      }
      pos += 33;
      continue;
    }
    if ((hdr >> 21) & 1u) {
      // block-cooperative record: every warp holds it at the same stream position
      const uint32_t n0 = pc[1 + nargs], n1 = pc[2 + nargs];
      if (op == trx::OP_X_DENSE && n0 <= 128)
        dense_forward_mma(pc + 1, m, dense_x);
      else if (op == trx::OP_R_X_DENSE && n0 <= 128 && n1 <= 128)
        dense_reverse_mma(pc + 1, m, dense_x, dense_g,
                          (kTrace && trace && blockIdx.x == 0) ? trace + (1ll << 20) + 32 * 4096 * 4 : nullptr);
      else
        trxvm::exec_coop(op, pc + 1, m, (int)threadIdx.x, (int)blockDim.x);
      pos += 1 + nargs + 2;
      continue;
    }
    }
    if (rec_trace) asm volatile("{ .reg .u32 d; add.u32 d, %1, 0; mov.u64 %0, %%clock64; }" : "=l"(rt1) : "r"(hdr));
    // the record layout is a compile-time property of each branch: operand addresses become
    // base + immediate, and exactly the operator's argument words are read from the staged page
    if (wide) {
      trxvm::CtxT<32, 1, true> c;
      c.m = m;
      c.nargs = nargs;
      c.preloaded = false;
      c.a = pc + 1 + lane32;
      if (lane32 < nsub) trxvm::exec<kCollision>(op, c);
      pos += 1 + nargs * 32;
    } else {
      trxvm::CtxT<4, 8, true> c;
      c.m = m;
      c.nargs = nargs;
      c.preloaded = false;
      c.lo = lane32 & 7;
      c.a = pc + 1 + (lane32 >> 3);
      if ((lane32 >> 3) < nsub) trxvm::exec<kCollision>(op, c);
      pos += 1 + nargs * 4;
    }
    if (rec_trace) {
      // development aid: per-record {start, header decoded, loads done, opcode}
      __syncwarp();
      __threadfence_block();
      long long rt2 = clock64();
      if (lane32 == 0) {
        long long *rec = trace + (1ll << 20) + ((long long)warp * 4096 + (rec_count & 4095)) * 4;
        rec[0] = rt0;
        rec[1] = rt1;
        rec[2] = rt2;
        rec[3] = (long long)op | ((long long)nsub << 16) | ((long long)trace_level << 32);
      }
      rec_count++;
    }
  }
}

// Constants of a program into every tile (simple.cpp:27-30).
__global__ void trx_const_kernel(const uint32_t *__restrict__ idx, const uint64_t *__restrict__ bits, uint32_t n,
                                 double *mem, uint64_t tile_stride) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t *m = reinterpret_cast<uint64_t *>(mem + (uint64_t)blockIdx.y * tile_stride);
  m[idx[i]] = bits[i];
}

struct DevElem {
  uint32_t idx;
  uint32_t fixed;
  uint32_t per_lane;
  uint32_t lane;  // 0xff = scalar-typed
};

// wide packed buffer -> tile images (base.cpp:9-31).  scalar_once: scalar-typed values go to tile 0
// only and the other tiles get zero (reverse-mode programs).
// `lanes` is the lane count of the wide buffer, `lane_base` the first lane handled by tile 0 of this
// memory (non-zero when an objective walks its rollouts in chunks).
__global__ void trx_scatter_kernel(const DevElem *__restrict__ elems, uint32_t n, const double *__restrict__ wide,
                                   uint64_t lanes, uint64_t lane_base, double *mem, uint64_t tile_stride,
                                   int scalar_once) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t tile = blockIdx.y;
  DevElem e = elems[i];
  double v;
  if (e.lane == 0xff) {
    v = wide[(uint64_t)e.fixed + (uint64_t)e.per_lane * lanes];
    if (scalar_once && (tile != 0 || lane_base != 0)) v = 0.0;
  } else {
    v = wide[(uint64_t)e.fixed + (uint64_t)e.per_lane * lanes + lane_base + tile * 8 + e.lane];
  }
  mem[tile * tile_stride + e.idx] = v;
}

// tile images -> wide packed buffer (base.cpp:33-40).  scalar_sum: scalar-typed values are summed
// over the tiles in tile order (deterministic); otherwise they are read from tile 0.
// accumulate: scalar-typed sums are added to what `wide` already holds (chunked objectives).
__global__ void trx_gather_kernel(const DevElem *__restrict__ elems, uint32_t n, double *__restrict__ wide,
                                  uint64_t lanes, uint64_t lane_base, const double *mem, uint64_t tile_stride,
                                  uint32_t n_tiles, int scalar_sum, int accumulate) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t tile = blockIdx.y;
  DevElem e = elems[i];
  if (e.lane == 0xff) {
    if (tile != 0) return;
    double v = mem[e.idx];
    if (scalar_sum)
      for (uint32_t t = 1; t < n_tiles; t++) v += mem[(uint64_t)t * tile_stride + e.idx];
    double *dst = wide + (uint64_t)e.fixed + (uint64_t)e.per_lane * lanes;
    *dst = accumulate ? (*dst + v) : v;
  } else {
    wide[(uint64_t)e.fixed + (uint64_t)e.per_lane * lanes + lane_base + tile * 8 + e.lane] =
        mem[tile * tile_stride + e.idx];
  }
}

// Objective glue (solvers/gd.h:50-63): r = outputs of the forward tape stay in the tile image;
// this kernel seeds the adjoint sweep with them (gradient slot = forward slot + prep.memory_size,
// gradients.cpp:80-86) and accumulates loss = |r|^2 per tile.  Scalar-typed residual rows are
// identical in every tile and count once.
struct SeedElem {
  uint32_t src, dst;
  uint32_t scalar;
};
constexpr int kSeedSplit = 8;  // blocks per tile (grid.y); partial sums are combined in a fixed order
__global__ void trx_seed_kernel(const SeedElem *__restrict__ elems, uint32_t n, double *mem, uint64_t tile_stride,
                                int first_chunk, double *__restrict__ partial) {
  __shared__ double red[256];
  double *m = mem + (uint64_t)blockIdx.x * tile_stride;
  bool owner = first_chunk && blockIdx.x == 0;
  const uint32_t per = (n + gridDim.y - 1) / gridDim.y;
  const uint32_t begin = blockIdx.y * per, end = min(n, begin + per);
  double acc = 0.0;
  uint32_t i = begin + threadIdx.x;
  // four independent element chains in flight per thread (the table and the residual rows are both
  // read once: latency, not reuse, is what limits this kernel)
  for (; i + 3 * blockDim.x < end; i += 4 * blockDim.x) {
    SeedElem e0 = elems[i], e1 = elems[i + blockDim.x], e2 = elems[i + 2 * blockDim.x], e3 = elems[i + 3 * blockDim.x];
    double v0 = m[e0.src], v1 = m[e1.src], v2 = m[e2.src], v3 = m[e3.src];
    if (e0.scalar && !owner) v0 = 0.0;
    if (e1.scalar && !owner) v1 = 0.0;
    if (e2.scalar && !owner) v2 = 0.0;
    if (e3.scalar && !owner) v3 = 0.0;
    m[e0.dst] = v0;
    m[e1.dst] = v1;
    m[e2.dst] = v2;
    m[e3.dst] = v3;
    acc += v0 * v0;
    acc += v1 * v1;
    acc += v2 * v2;
    acc += v3 * v3;
  }
  for (; i < end; i += blockDim.x) {
    SeedElem e = elems[i];
    double v = m[e.src];
    if (e.scalar && !owner) v = 0.0;
    m[e.dst] = v;
    acc += v * v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)blockIdx.x * gridDim.y + blockIdx.y] = red[0];
}
// fixed-order sum of the per-tile partials: out[0] (+)= sum
__global__ void trx_loss_kernel(const double *__restrict__ partial, uint32_t n, double *out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// GradientDescentSolver::_step after the engine work (solvers/gd.h:64-129): v <- mu v - lr g ; x <- x + v,
// skipped altogether when any gradient element is not finite (gd.h:72,124-126).  One CTA: the optimisation
// vector is the policy's few thousand weights.  out = {loss, g[n]} of trx_objective_evaluate_device.
__global__ void trx_gd_update_kernel(double *x, double *v, const double *out, uint32_t n, double lr, double mu,
                                     double *loss_log) {
  const double *g = out + 1;
  int finite = 1;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) finite &= isfinite(g[i]) ? 1 : 0;
  finite = __syncthreads_and(finite);
  if (threadIdx.x == 0 && loss_log) *loss_log = out[0];
  if (!finite) return;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double vi = mu * v[i] - lr * g[i];
    v[i] = vi;
    x[i] = x[i] + vi;
  }
}

// ---- Gauss-Newton pieces (solvers/sq.h, solvers/base.h:76-85)
// after the tangent sweep: scalar-typed (lane-independent) output tangents are the same in every tile and seed the
// adjoint once, in tile 0 of the rank that owns those rows (the same rule as trx_seed_kernel)
__global__ void trx_mask_scalar_kernel(const DevElem *__restrict__ elems, uint32_t n, double *mem, uint64_t tile_stride,
                                       int keep_in_tile0) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t tile = blockIdx.y;
  const DevElem e = elems[i];
  if (e.lane == 0xff && (tile != 0 || !keep_in_tile0)) mem[tile * tile_stride + e.idx] = 0.0;
}
// out[slot] = a . b ; one CTA, fixed summation order (deterministic)
__global__ void trx_dot_kernel(const double *__restrict__ a, const double *__restrict__ b, uint32_t n, double *out) {
  __shared__ double red[1024];
  double sum = 0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sum += a[i] * b[i];
  red[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t w = blockDim.x / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}
// y <- y + alpha x   (and, when z is given, z <- z + beta x)
__global__ void trx_axpy2_kernel(double *y, double alpha, const double *__restrict__ x, double *z, double beta, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xi = x[i];
  y[i] += alpha * xi;
  if (z) z[i] += beta * xi;
}
// p <- r + beta p
__global__ void trx_cg_direction_kernel(double *p, const double *__restrict__ r, double beta, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = r[i] + beta * p[i];
}

}  // namespace

// ------------------------------------------------------------------ objects

struct trx_program {
  int device = 0;
  trx::LoweredProgram low;
  uint32_t *d_stream = nullptr;
  uint64_t *d_stream_begin = nullptr;
  uint32_t *d_const_idx = nullptr;
  uint64_t *d_const_bits = nullptr;
  double *d_pool = nullptr;  // constant pool of the on-chip slot cache (dummy allocation when empty)
  DevElem *d_elems[3] = {nullptr, nullptr, nullptr};
  uint32_t n_elems[3] = {0, 0, 0};
  int scalar_inputs = TRX_SCALAR_SHARED, scalar_outputs = TRX_SCALAR_SHARED;
  bool has_scalar_outputs = false;  // any output port of scalar type (one value for all lanes)
  uint64_t shape_generation = 0;    // collision shape table the program's handles were checked against
  const trx::PortLayout &layout(int which) const {
    return which == TRX_BUF_INPUT ? low.inputs : which == TRX_BUF_PARAMETER ? low.parameters : low.outputs;
  }
};

struct trx_memory {
  int device = 0;
  uint64_t lanes = 0, n_tiles = 0;
  uint64_t tile_stride = 0;  // doubles
  double *d_mem = nullptr;
  double *d_stage = nullptr;
  uint64_t stage_elems = 0;
  double *h_stage = nullptr;  // pinned
  uint64_t h_stage_elems = 0;
  uint64_t rng_lane_base = 0;  // rollout of lane 0 of tile 0 (objectives that walk their rollouts in chunks)
};

namespace {

trx_status ensureTileStride(trx_memory *m, uint64_t bytes_per_tile, cudaStream_t stream) {
  uint64_t need = (bytes_per_tile + 255) / 256 * 256 / 8;
  if (need <= m->tile_stride) return TRX_OK;
  double *fresh = nullptr;
  TRX_CUDA(cudaMalloc(&fresh, need * m->n_tiles * sizeof(double)));
  TRX_CUDA(cudaMemsetAsync(fresh, 0, need * m->n_tiles * sizeof(double), stream));
  if (m->d_mem) {
    // keep what earlier programs left in the image (MemoryImpl::resize keeps contents, base.h:24)
    TRX_CUDA(cudaMemcpy2DAsync(fresh, need * sizeof(double), m->d_mem, m->tile_stride * sizeof(double),
                               m->tile_stride * sizeof(double), m->n_tiles, cudaMemcpyDeviceToDevice, stream));
    TRX_CUDA(cudaStreamSynchronize(stream));
    TRX_CUDA(cudaFree(m->d_mem));
  }
  m->d_mem = fresh;
  m->tile_stride = need;
  return TRX_OK;
}

trx_status ensureStage(trx_memory *m, uint64_t elems, bool host) {
  if (elems > m->stage_elems) {
    if (m->d_stage) TRX_CUDA(cudaFree(m->d_stage));
    m->d_stage = nullptr;
    TRX_CUDA(cudaMalloc(&m->d_stage, elems * sizeof(double)));
    m->stage_elems = elems;
  }
  if (host && elems > m->h_stage_elems) {
    if (m->h_stage) TRX_CUDA(cudaFreeHost(m->h_stage));
    m->h_stage = nullptr;
    TRX_CUDA(cudaMallocHost(&m->h_stage, elems * sizeof(double)));
    m->h_stage_elems = elems;
  }
  return TRX_OK;
}

template <class T> trx_status upload(const std::vector<T> &v, T **dst) {
  *dst = nullptr;
  if (v.empty()) return TRX_OK;
  TRX_CUDA(cudaMalloc(dst, v.size() * sizeof(T)));
  TRX_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return TRX_OK;
}

trx_status scatterDevice(const trx_program *p, trx_memory *m, int which, const double *d_wide, uint64_t bytes,
                         cudaStream_t stream, uint64_t wide_lanes = 0, uint64_t lane_base = 0) {
  const trx::PortLayout &L = p->layout(which);
  if (wide_lanes == 0) wide_lanes = m->lanes;
  uint64_t expect = L.wideElems(wide_lanes) * 8;
  if (bytes != expect)
    return fail(TRX_ERR_SIZE, "packed buffer size mismatch: got " + std::to_string(bytes) + " expected " +
                                  std::to_string(expect));
  trx_status s = ensureTileStride(m, p->low.memory_size, stream);
  if (s != TRX_OK) return s;
  uint32_t n = p->n_elems[which];
  if (n == 0) return TRX_OK;
  dim3 grid((n + 255) / 256, (unsigned)m->n_tiles);
  int once = (which == TRX_BUF_INPUT && p->scalar_inputs == TRX_SCALAR_REDUCE) ? 1 : 0;
  trx_scatter_kernel<<<grid, 256, 0, stream>>>(p->d_elems[which], n, d_wide, wide_lanes, lane_base, m->d_mem,
                                               m->tile_stride, once);
  g_launches++;
  TRX_CUDA(cudaGetLastError());
  return TRX_OK;
}

trx_status scatterHost(const trx_program *p, trx_memory *m, int which, const void *packed, uint64_t bytes,
                       cudaStream_t stream) {
  const trx::PortLayout &L = p->layout(which);
  uint64_t elems = L.wideElems(m->lanes);
  if (bytes != elems * 8)
    return fail(TRX_ERR_SIZE, "packed buffer size mismatch: got " + std::to_string(bytes) + " expected " +
                                  std::to_string(elems * 8));
  if (elems == 0) return TRX_OK;
  // the staging buffers are reused: earlier work on this stream must be done with them
  TRX_CUDA(cudaStreamSynchronize(stream));
  trx_status s = ensureStage(m, elems, true);
  if (s != TRX_OK) return s;
  std::memcpy(m->h_stage, packed, bytes);
  TRX_CUDA(cudaMemcpyAsync(m->d_stage, m->h_stage, bytes, cudaMemcpyHostToDevice, stream));
  return scatterDevice(p, m, which, m->d_stage, bytes, stream);
}

}  // namespace

// ------------------------------------------------------------------ C ABI

extern "C" {

const char *trx_last_error(void) { return g_last_error.c_str(); }
const char *trx_version(void) { return "trx-b200 0.1 (sm_100a)"; }

int trx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int trx_op_count(void) { return (int)trx::opTable().size(); }
const char *trx_op_name(int i) {
  auto &t = trx::opTable();
  return (i >= 0 && i < (int)t.size()) ? t[i].tape_name.c_str() : nullptr;
}
const char *trx_op_signature(int i) {
  auto &t = trx::opTable();
  return (i >= 0 && i < (int)t.size()) ? t[i].sig : nullptr;
}

uint64_t trx_kernel_launch_count(void) { return g_launches.load(); }

// development aid: per-level bundle counts of a generically lowered program (low 16 bits: bundles,
// high 16 bits: block-cooperative records)
trx_status trx_debug_level_info(const trx_program *p, uint32_t *out, uint64_t n) {
  if (!p || !out) return fail(TRX_ERR_INVALID, "bad argument");
  for (uint64_t i = 0; i < n && i < p->low.level_bundles.size(); i++) out[i] = p->low.level_bundles[i];
  return TRX_OK;
}

// development aid: the next generic-VM launches record, for tile 0, the clock at which every warp
// arrives at / leaves each level barrier.  n_words = capacity of the device buffer in 64-bit words.
trx_status trx_debug_trace(uint64_t n_words, long long **device_buffer) {
  if (g_trace) cudaFree(g_trace);
  g_trace = nullptr;
  if (n_words) {
    TRX_CUDA(cudaMalloc(&g_trace, n_words * 8));
    TRX_CUDA(cudaMemset(g_trace, 0, n_words * 8));
  }
  if (device_buffer) *device_buffer = g_trace;
  return TRX_OK;
}

// uploads p->low (already lowered) to the device
static trx_status finishProgram(std::unique_ptr<trx_program> &p, int device, int scalar_in, int scalar_out,
                                trx_program **out);

static trx_status checkShapeHandles(const trxfile::Program &src, uint64_t *generation);
static trx_status createFromFile(const trxfile::Program &src, int device, int scalar_in, int scalar_out,
                                 trx_program **out, bool fold_accumulate = false) {
  if (!out) return fail(TRX_ERR_INVALID, "null output pointer");
  *out = nullptr;
  std::unique_ptr<trx_program> p(new trx_program());
  p->device = device;
  try {
    trx::LowerOptions opt;
    if (const char *w = getenv("TRX_PREFETCH")) opt.prefetch_distance = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_CACHE_SLOTS")) opt.ring_slots = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_POOL_SLOTS")) opt.pool_slots = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_CHAIN")) opt.chain_cap = (uint32_t)atoi(w);
    // the staged kernel runs 20 warps per tile (640 threads x <= 96 registers fill the register file,
    // 20 x 8 KB of stream pages + 16 KB of dense staging fit shared memory); the other kernels 16
    const bool staged = staged_streams() && opt.ring_slots == 0 && opt.pool_slots == 0 && opt.page_words == 512;
    opt.n_warps = staged ? kStagedWarps : 16;
    if (const char *w = getenv("TRX_WARPS")) opt.n_warps = (uint32_t)atoi(w);
    if (opt.n_warps > (staged ? (uint32_t)kStagedWarps : 16u)) return fail(TRX_ERR_INVALID, "TRX_WARPS too large for this kernel");
    // folding needs the plain argument-word format of the staged kernel
    static int fold_env = getenv("TRX_FOLD") ? atoi(getenv("TRX_FOLD")) : 1;
    opt.fold_accumulate = fold_accumulate && fold_env && staged_streams() && opt.ring_slots == 0 && opt.pool_slots == 0 &&
                          opt.page_words == 512;
    trx::lowerProgram(src, opt, p->low);
  } catch (const std::exception &e) {
    std::string msg = e.what();
    bool unsupported = msg.find("unsupported operator") != std::string::npos;
    return fail(unsupported ? TRX_ERR_UNSUPPORTED : TRX_ERR_INVALID, msg);
  }
  if (p->low.uses_shapes) {
    trx_status cs = checkShapeHandles(src, &p->shape_generation);
    if (cs != TRX_OK) return cs;
  }
  return finishProgram(p, device, scalar_in, scalar_out, out);
}

static trx_status finishProgram(std::unique_ptr<trx_program> &p, int device, int scalar_in, int scalar_out,
                                trx_program **out) {
  // scalar-typed port behaviour across tiles (see trx_b200.h)
  bool reverse_only = p->low.has_reverse && !p->low.has_forward;
  p->has_scalar_outputs = p->low.outputs.fixed_total > 0;
  p->scalar_inputs = scalar_in != TRX_SCALAR_AUTO ? scalar_in : (reverse_only ? TRX_SCALAR_REDUCE : TRX_SCALAR_SHARED);
  p->scalar_outputs = scalar_out != TRX_SCALAR_AUTO ? scalar_out : (p->low.has_reverse ? TRX_SCALAR_REDUCE : TRX_SCALAR_SHARED);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(TRX_ERR_CUDA, "no CUDA device: the tape engine has no CPU execution path");
  TRX_CUDA(cudaSetDevice(device));
  trx_status s;
  if ((s = upload(p->low.stream, &p->d_stream)) != TRX_OK) return s;
  if ((s = upload(p->low.stream_begin, &p->d_stream_begin)) != TRX_OK) return s;
  if ((s = upload(p->low.const_idx, &p->d_const_idx)) != TRX_OK) return s;
  if ((s = upload(p->low.const_bits, &p->d_const_bits)) != TRX_OK) return s;
  if (p->low.cta_ring || p->low.global_pool) {
    std::vector<double> pool = p->low.pool_values;
    if (pool.empty()) pool.push_back(0.0);
    if ((s = upload(pool, &p->d_pool)) != TRX_OK) return s;
  }
  for (int which = 0; which < 3; which++) {
    const trx::PortLayout &L = p->layout(which);
    std::vector<DevElem> elems(L.elems.size());
    for (size_t i = 0; i < elems.size(); i++)
      elems[i] = {L.elems[i].idx, L.elems[i].fixed, L.elems[i].per_lane, L.elems[i].lane};
    p->n_elems[which] = (uint32_t)elems.size();
    if ((s = upload(elems, &p->d_elems[which])) != TRX_OK) return s;
  }
  *out = p.release();
  return TRX_OK;
}

trx_status trx_program_create(const trx_program_desc *d, int device, trx_program **out) {
  if (!d) return fail(TRX_ERR_INVALID, "null program description");
  trxfile::Program src;
  src.memory_size = d->memory_size;
  src.ops.resize(d->n_ops);
  for (uint32_t i = 0; i < d->n_ops; i++) {
    src.ops[i].name = d->ops[i].name ? d->ops[i].name : "";
    src.ops[i].args.resize(d->ops[i].n_args);
    for (uint32_t k = 0; k < d->ops[i].n_args; k++) {
      src.ops[i].args[k].size = d->ops[i].args[k].size;
      src.ops[i].args[k].is_input = d->ops[i].args[k].is_input;
      src.ops[i].args[k].lanes = d->ops[i].args[k].lanes;
      src.ops[i].args[k].elem = d->ops[i].args[k].elem;
    }
  }
  src.n_instructions = d->n_instructions;
  src.code.assign(d->code, d->code + d->n_code_words);
  auto ports = [](const trx_port_desc *p, uint32_t n, std::vector<trxfile::Port> &dst) {
    dst.resize(n);
    for (uint32_t i = 0; i < n; i++) {
      dst[i].address = p[i].address;
      dst[i].offset = p[i].offset;
      dst[i].size = p[i].size;
      dst[i].lanes = p[i].lanes;
      dst[i].elem = p[i].elem;
    }
  };
  ports(d->inputs, d->n_inputs, src.inputs);
  ports(d->parameters, d->n_parameters, src.parameters);
  ports(d->outputs, d->n_outputs, src.outputs);
  ports(d->constants, d->n_constants, src.constants);
  src.const_data.assign(d->const_data, d->const_data + d->n_const_data);
  return createFromFile(src, device, d->scalar_inputs, d->scalar_outputs, out);
}

trx_status trx_program_create_from_blob(const void *blob, size_t bytes, int device, trx_program **out) {
  trxfile::Program src;
  try {
    trxfile::decodeProgram(blob, bytes, src);
  } catch (const std::exception &e) {
    return fail(TRX_ERR_INVALID, e.what());
  }
  return createFromFile(src, device, TRX_SCALAR_AUTO, TRX_SCALAR_AUTO, out);
}

void trx_program_destroy(trx_program *p) {
  if (!p) return;
  cudaFree(p->d_stream);
  cudaFree(p->d_stream_begin);
  cudaFree(p->d_const_idx);
  cudaFree(p->d_const_bits);
  cudaFree(p->d_pool);
  for (auto *e : p->d_elems) cudaFree(e);
  delete p;
}

trx_status trx_program_buffer_bytes(const trx_program *p, int which, uint64_t lanes, uint64_t *bytes) {
  if (!p || !bytes || which < 0 || which > 2) return fail(TRX_ERR_INVALID, "bad argument");
  *bytes = p->layout(which).wideElems(lanes) * 8;
  return TRX_OK;
}

trx_status trx_program_stat(const trx_program *p, const char *key, uint64_t *value) {
  if (!p || !key || !value) return fail(TRX_ERR_INVALID, "bad argument");
  std::string k = key;
  if (k == "n_instructions") *value = p->low.n_instructions;
  else if (k == "n_levels") *value = p->low.n_levels;
  else if (k == "n_bundles") *value = p->low.n_bundles;
  else if (k == "stream_words") *value = p->low.stream.size();
  else if (k == "memory_size") *value = p->low.memory_size;
  else if (k == "n_warps") *value = p->low.n_warps;
  else if (k == "arg_bytes_per_tile") *value = p->low.arg_bytes_per_tile;
  else if (k == "max_level_width") *value = p->low.max_level_width;
  else if (k == "n_constants") *value = p->low.const_idx.size();
  else if (k == "ring_doubles") *value = p->low.ring_doubles;
  else if (k == "n_prefetch_lines") *value = p->low.n_prefetch_lines;
  else if (k == "n_chained") *value = p->low.n_chained;
  else if (k == "n_folded") *value = p->low.n_folded;
  else if (k == "n_ring_reads") *value = p->low.n_ring_reads;
  else if (k == "n_image_reads") *value = p->low.n_image_reads;
  else if (k == "n_ring_only_writes") *value = p->low.n_ring_only_writes;
  else if (k == "n_dual_writes") *value = p->low.n_dual_writes;
  else if (k == "n_image_writes") *value = p->low.n_image_writes;
  else if (k == "scalar_inputs") *value = (uint64_t)p->scalar_inputs;
  else if (k == "scalar_outputs") *value = (uint64_t)p->scalar_outputs;
  else return fail(TRX_ERR_INVALID, "unknown statistic " + k);
  return TRX_OK;
}

trx_status trx_memory_create(uint64_t lanes, int device, trx_memory **out) {
  if (!out) return fail(TRX_ERR_INVALID, "null output pointer");
  *out = nullptr;
  if (lanes == 0 || lanes % 8 != 0) return fail(TRX_ERR_INVALID, "lanes must be a positive multiple of 8");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(TRX_ERR_CUDA, "no CUDA device: the tape engine has no CPU execution path");
  TRX_CUDA(cudaSetDevice(device));
  trx_memory *m = new trx_memory();
  m->device = device;
  m->lanes = lanes;
  m->n_tiles = lanes / 8;
  *out = m;
  return TRX_OK;
}

void trx_memory_destroy(trx_memory *m) {
  if (!m) return;
  cudaFree(m->d_mem);
  cudaFree(m->d_stage);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  delete m;
}

trx_status trx_memory_lanes(const trx_memory *m, uint64_t *lanes) {
  if (!m || !lanes) return fail(TRX_ERR_INVALID, "bad argument");
  *lanes = m->lanes;
  return TRX_OK;
}

trx_status trx_memory_bytes(const trx_memory *m, uint64_t *bytes) {
  if (!m || !bytes) return fail(TRX_ERR_INVALID, "bad argument");
  *bytes = m->tile_stride * m->n_tiles * 8;
  return TRX_OK;
}

trx_status trx_input(const trx_program *p, trx_memory *m, const void *packed, uint64_t bytes, void *stream) {
  if (!p || !m || (!packed && bytes)) return fail(TRX_ERR_INVALID, "bad argument");
  TRX_CUDA(cudaSetDevice(m->device));
  return scatterHost(p, m, TRX_BUF_INPUT, packed, bytes, (cudaStream_t)stream);
}
trx_status trx_parameterize(const trx_program *p, trx_memory *m, const void *packed, uint64_t bytes, void *stream) {
  if (!p || !m || (!packed && bytes)) return fail(TRX_ERR_INVALID, "bad argument");
  TRX_CUDA(cudaSetDevice(m->device));
  return scatterHost(p, m, TRX_BUF_PARAMETER, packed, bytes, (cudaStream_t)stream);
}
trx_status trx_input_device(const trx_program *p, trx_memory *m, const void *d, uint64_t bytes, void *stream) {
  if (!p || !m || (!d && bytes)) return fail(TRX_ERR_INVALID, "bad argument");
  TRX_CUDA(cudaSetDevice(m->device));
  return scatterDevice(p, m, TRX_BUF_INPUT, (const double *)d, bytes, (cudaStream_t)stream);
}
trx_status trx_parameterize_device(const trx_program *p, trx_memory *m, const void *d, uint64_t bytes, void *stream) {
  if (!p || !m || (!d && bytes)) return fail(TRX_ERR_INVALID, "bad argument");
  TRX_CUDA(cudaSetDevice(m->device));
  return scatterDevice(p, m, TRX_BUF_PARAMETER, (const double *)d, bytes, (cudaStream_t)stream);
}

static trx_status executeImpl(const trx_program *p, trx_memory *m, void *stream_, bool upload_constants);
trx_status trx_execute(const trx_program *p, trx_memory *m, void *stream_) { return executeImpl(p, m, stream_, true); }

// upload_constants = false: the caller knows that this memory already holds the program's constants and
// that nothing has overwritten them (objectives, after their first evaluation)
static std::atomic<uint64_t> g_rng_seed{0x9E3779B97F4A7C15ull}, g_rng_epoch{0};

// ---- collision shape table (include/trx_b200.h, csrc/collision/trx_collide.h): one host copy, one mirror per device
struct ShapeRegistry {
  std::mutex mu;
  trxcol::ShapeTable table;
  uint64_t version = 1;
  uint64_t generation = 1;  // bumped by trx_collision_clear: handles of an earlier generation are dead
  struct Mirror {
    double *d = nullptr;
    uint64_t version = 0, capacity = 0;
  } mirror[64];
};
static ShapeRegistry &shapeRegistry() {
  static ShapeRegistry r;
  return r;
}
// makes the device's copy current and points the kernels' symbol at it (ordered on `stream`)
static trx_status syncShapeTable(int device, cudaStream_t stream) {
  ShapeRegistry &R = shapeRegistry();
  std::lock_guard<std::mutex> lock(R.mu);
  if (device < 0 || device >= 64) return fail(TRX_ERR_INVALID, "device index out of range");
  auto &M = R.mirror[device];
  if (M.version == R.version) return TRX_OK;
  if (M.capacity < R.table.heap.size()) {
    // (an old table may still be read by work in flight: it is left to the process' end rather than freed)
    M.capacity = std::max<uint64_t>(2 * R.table.heap.size(), 1024);
    TRX_CUDA(cudaMalloc(&M.d, M.capacity * sizeof(double)));
  }
  TRX_CUDA(cudaMemcpyAsync(M.d, R.table.heap.data(), R.table.heap.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  TRX_CUDA(cudaStreamSynchronize(stream));  // the host vector may grow once the lock is released
  const double *ptr = M.d;
  TRX_CUDA(cudaMemcpyToSymbolAsync(trxvm::g_shape_heap_dev, &ptr, sizeof(ptr), 0, cudaMemcpyHostToDevice, stream));
  M.version = R.version;
  return TRX_OK;
}
// every uint64 constant that a collision instruction reads must be a registered handle of the right kind
static trx_status checkShapeHandles(const trxfile::Program &src, uint64_t *generation) {
  ShapeRegistry &R = shapeRegistry();
  std::lock_guard<std::mutex> lock(R.mu);
  *generation = R.generation;
  std::unordered_map<uint64_t, uint64_t> const_at;  // address -> bits
  for (auto &c : src.constants)
    if (c.size == 8 && c.offset + 8 <= src.const_data.size()) {
      uint64_t bits;
      memcpy(&bits, src.const_data.data() + c.offset, 8);
      const_at[c.address] = bits;
    }
  uint64_t pc = 0;
  while (pc < src.code.size()) {  // (lowerProgram has validated the code words)
    const auto &op = src.ops[src.code[pc]];
    const bool axes = op.name.rfind("collision_axes_", 0) == 0, proj = op.name.rfind("collision_project_2_", 0) == 0;
    if (axes || proj) {
      auto it = const_at.find(src.code[pc + 1 + (axes ? 2 : 1)]);
      if (it == const_at.end())
        return fail(TRX_ERR_INVALID, op.name + ": the shape argument is not a constant of the program");
      const auto &known = axes ? R.table.pairs : R.table.shapes;
      if (std::find(known.begin(), known.end(), it->second) == known.end())
        return fail(TRX_ERR_INVALID, op.name + ": constant " + std::to_string(it->second) +
                                         " is not a handle from trx_collision_" + (axes ? "pair" : "shape") + "_create");
    }
    pc += 1 + op.args.size();
  }
  return TRX_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember which devices have it
static bool firstUseOnDevice(int device, int which) {
  static std::mutex mu;
  static uint64_t done[2] = {0, 0};  // bit = device
  std::lock_guard<std::mutex> lock(mu);
  if (device < 0 || device >= 64) return true;
  const bool first = !(done[which] >> device & 1);
  done[which] |= 1ull << device;
  return first;
}

static trx_status executeImpl(const trx_program *p, trx_memory *m, void *stream_, bool upload_constants) {
  if (!p || !m) return fail(TRX_ERR_INVALID, "bad argument");
  if (p->device != m->device) return fail(TRX_ERR_INVALID, "executable and memory live on different devices");
  // a program with forward- and reverse-mode instructions (hprop) propagates scalar-typed (lane-independent)
  // output tangents into the adjoint in every lane tile, so they would count once per tile: refuse instead of
  // returning a wrong J^T J v (one tile is exact; trx_objective_hessian_product handles any number of tiles)
  if (p->low.has_forward && p->low.has_reverse && m->n_tiles > 1 && p->has_scalar_outputs)
    return fail(TRX_ERR_UNSUPPORTED,
                "forward+reverse program with scalar-typed outputs on more than one lane tile: run it on 8 lanes, or "
                "use trx_objective_hessian_product");
  cudaStream_t stream = (cudaStream_t)stream_;
  TRX_CUDA(cudaSetDevice(m->device));
  trx_status s = ensureTileStride(m, p->low.memory_size, stream);
  if (s != TRX_OK) return s;
  if (p->low.draws_random) {
    // in-tape random draws (vm_exec.cuh): fresh values at every execute, reproducible from the seed
    trxvm::RngState st;
    st.seed = g_rng_seed.load();
    st.epoch = ++g_rng_epoch;
    st.base = m->d_mem;
    st.tile_stride = m->tile_stride;
    st.lane_base = m->rng_lane_base;
    TRX_CUDA(cudaMemcpyToSymbolAsync(trxvm::g_rng_dev, &st, sizeof(st), 0, cudaMemcpyHostToDevice, stream));
  }
  if (p->low.uses_shapes) {
    {
      ShapeRegistry &R = shapeRegistry();
      std::lock_guard<std::mutex> lock(R.mu);
      if (p->shape_generation != R.generation)
        return fail(TRX_ERR_INVALID, "the collision shape table was cleared after this program was created: its "
                                     "shape handles are dead");
    }
    trx_status cs = syncShapeTable(m->device, stream);
    if (cs != TRX_OK) return cs;
  }
  uint32_t nc = upload_constants ? (uint32_t)p->low.const_idx.size() : 0;
  if (nc) {
    dim3 grid((nc + 255) / 256, (unsigned)m->n_tiles);
    trx_const_kernel<<<grid, 256, 0, stream>>>(p->d_const_idx, p->d_const_bits, nc, m->d_mem, m->tile_stride);
    g_launches++;
  }
  if (p->low.n_instructions) {
    uint32_t W = p->low.n_warps;
    size_t smem = (size_t)W * p->low.ring_doubles * sizeof(double);
    if (p->low.ring_doubles) {
#ifndef TRX_WITH_EXPERIMENTS
      return fail(TRX_ERR_UNSUPPORTED, "on-chip slot rings are an experiment that is not built into this library");
#else
      if (firstUseOnDevice(m->device, 0))
        TRX_CUDA(cudaFuncSetAttribute(trx_vm_kernel<16, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      if (p->low.cta_ring) smem = (size_t)p->low.ring_doubles * sizeof(double);
      trx_vm_kernel<16, true, true, false><<<(unsigned)m->n_tiles, W * 32, smem, stream>>>(
          p->d_stream, p->d_stream_begin, m->d_mem, m->tile_stride, p->low.ring_doubles, g_trace,
          p->low.cta_ring ? p->d_pool : nullptr, p->low.pool_offset_doubles, (uint32_t)p->low.pool_values.size(),
          p->low.page_words);
#endif
    } else if (p->low.page_words == 512 && !p->low.global_pool && staged_streams()) {
      // the product path: stream pages staged in shared memory
      constexpr int kPages = 4;
      const size_t bytes = (size_t)kStagedWarps * 512 * kPages * sizeof(uint32_t);
      if (firstUseOnDevice(m->device, 1)) {
        TRX_CUDA(cudaFuncSetAttribute(trx_vm_staged_kernel<kStagedWarps, 512, kPages, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        TRX_CUDA(cudaFuncSetAttribute(trx_vm_staged_kernel<kStagedWarps, 512, kPages, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        TRX_CUDA(cudaFuncSetAttribute(trx_vm_staged_kernel<kStagedWarps, 512, kPages, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      }
      const size_t smem_bytes = (size_t)W * 512 * kPages * sizeof(uint32_t);
      if (p->low.uses_shapes)  // the instantiation with the collision operators (GJK / EPA calls, 7 KB stack frame)
        trx_vm_staged_kernel<kStagedWarps, 512, kPages, false, true><<<(unsigned)m->n_tiles, W * 32, smem_bytes, stream>>>(
            p->d_stream, p->d_stream_begin, m->d_mem, m->tile_stride, nullptr);
      else if (g_trace)  // development aid (trx_debug_trace): the instrumented instantiation
        trx_vm_staged_kernel<kStagedWarps, 512, kPages, true, false><<<(unsigned)m->n_tiles, W * 32, smem_bytes, stream>>>(
            p->d_stream, p->d_stream_begin, m->d_mem, m->tile_stride, g_trace);
      else
        trx_vm_staged_kernel<kStagedWarps, 512, kPages, false, false><<<(unsigned)m->n_tiles, W * 32, smem_bytes, stream>>>(
            p->d_stream, p->d_stream_begin, m->d_mem, m->tile_stride, nullptr);
    } else {
      if (p->low.uses_shapes)
        return fail(TRX_ERR_UNSUPPORTED, "collision operators run in the staged interpreter only (TRX_STAGED=0 or a "
                                         "non-default page size selects a development kernel without them)");
      if (p->low.global_pool)
        trx_vm_kernel<16, false, true, true><<<(unsigned)m->n_tiles, W * 32, 0, stream>>>(
            p->d_stream, p->d_stream_begin, m->d_mem, m->tile_stride, 0, g_trace, p->d_pool, 0, 0, p->low.page_words);
      else
        trx_vm_kernel<16, false, true, false><<<(unsigned)m->n_tiles, W * 32, 0, stream>>>(
            p->d_stream, p->d_stream_begin, m->d_mem, m->tile_stride, 0, g_trace, nullptr, 0, 0, p->low.page_words);
    }
    g_launches++;
  }
  TRX_CUDA(cudaGetLastError());
  return TRX_OK;
}

trx_status trx_output_device(const trx_program *p, trx_memory *m, void *d_packed, uint64_t bytes, void *stream_) {
  if (!p || !m || (!d_packed && bytes)) return fail(TRX_ERR_INVALID, "bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  TRX_CUDA(cudaSetDevice(m->device));
  const trx::PortLayout &L = p->low.outputs;
  uint64_t expect = L.wideElems(m->lanes) * 8;
  if (bytes != expect)
    return fail(TRX_ERR_SIZE, "packed buffer size mismatch: got " + std::to_string(bytes) + " expected " +
                                  std::to_string(expect));
  if (m->tile_stride * 8 < p->low.memory_size)
    return fail(TRX_ERR_INVALID, "memory has not been written by a program of this size yet");
  uint32_t n = p->n_elems[TRX_BUF_OUTPUT];
  if (n == 0) return TRX_OK;
  dim3 grid((n + 255) / 256, (unsigned)m->n_tiles);
  trx_gather_kernel<<<grid, 256, 0, stream>>>(p->d_elems[TRX_BUF_OUTPUT], n, (double *)d_packed, m->lanes, 0, m->d_mem,
                                             m->tile_stride, (uint32_t)m->n_tiles,
                                             p->scalar_outputs == TRX_SCALAR_REDUCE ? 1 : 0, 0);
  g_launches++;
  TRX_CUDA(cudaGetLastError());
  return TRX_OK;
}

trx_status trx_output(const trx_program *p, trx_memory *m, void *packed, uint64_t bytes, void *stream_) {
  if (!p || !m || (!packed && bytes)) return fail(TRX_ERR_INVALID, "bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  TRX_CUDA(cudaSetDevice(m->device));
  uint64_t elems = p->low.outputs.wideElems(m->lanes);
  if (bytes != elems * 8)
    return fail(TRX_ERR_SIZE, "packed buffer size mismatch: got " + std::to_string(bytes) + " expected " +
                                  std::to_string(elems * 8));
  if (elems == 0) return TRX_OK;
  trx_status s = ensureStage(m, elems, true);
  if (s != TRX_OK) return s;
  s = trx_output_device(p, m, m->d_stage, bytes, stream_);
  if (s != TRX_OK) return s;
  TRX_CUDA(cudaMemcpyAsync(m->h_stage, m->d_stage, bytes, cudaMemcpyDeviceToHost, stream));
  TRX_CUDA(cudaStreamSynchronize(stream));
  std::memcpy(packed, m->h_stage, bytes);
  return TRX_OK;
}

trx_status trx_memory_peek(trx_memory *m, uint64_t tile, uint64_t address, void *dst, uint64_t bytes) {
  if (!m || !dst) return fail(TRX_ERR_INVALID, "bad argument");
  if (tile >= m->n_tiles || address + bytes > m->tile_stride * 8) return fail(TRX_ERR_INVALID, "peek out of range");
  TRX_CUDA(cudaSetDevice(m->device));
  TRX_CUDA(cudaDeviceSynchronize());
  TRX_CUDA(cudaMemcpy(dst, (const uint8_t *)m->d_mem + tile * m->tile_stride * 8 + address, bytes,
                      cudaMemcpyDeviceToHost));
  return TRX_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ objective (solver-level glue)

struct trx_objective {
  // re-rolled rollout objective (trx_rollout.cu): when set, the tape members below are unused
  trx_rollout *rollout = nullptr;
  // owned.  fused: prog = forward+prepare, prep = null, bprop = adjoint (trx_fuse.h);
  // otherwise the three programs lowered one by one as generic executables (trx_lower.h)
  trx_program *prog = nullptr, *prep = nullptr, *bprop = nullptr;
  bool fused = false;
  bool scalar_rows = true;  // false: lane-independent residual rows are another rank's to count
  int device = 0;
  uint64_t lanes = 0, chunk_lanes = 0, n_chunks = 0;
  uint64_t n_x = 0;
  trx_memory *mem = nullptr;
  double *d_x = nullptr, *d_params = nullptr, *d_out = nullptr, *d_partial = nullptr;
  trx::DirectStats direct;
  bool consts_static = false;    // no tape instruction or port of the objective overlaps a constant slot
  bool consts_resident = false;  // ... and the memory holds them since the first evaluation
  double *h_x = nullptr, *h_out = nullptr;  // pinned
  uint64_t param_elems = 0;
  SeedElem *d_seed = nullptr;
  uint32_t n_seed = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_ms[3] = {0, 0, 0};
  // optional per-program kernel timing (CUDA events on the launching stream)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;  // 4 per chunk: before prog, after prog, after prep(+seed), after bprop
  double prof_ms[3] = {0, 0, 0};
  uint64_t prof_count = 0;
  trx_comm *comm = nullptr;  // not owned: all-reduce of {loss, gradient} after every evaluation
  // Gauss-Newton (LeastSquaresSolver): tangent executable and the conjugate-gradient vectors, all on the device
  trx_program *fprop = nullptr;
  bool keeps_prep_state = false;  // x_prep's operand copies are all there (tangent rules read them)
  double *d_cg = nullptr;        // 5 x n_x: rhs, solution, residual, direction, product
  double *d_scal = nullptr;      // dot products
  double *h_scal = nullptr;      // pinned
};

extern "C" {

#ifdef TRX_WITH_EXPERIMENTS
// Fused lowering of the objective's two sweeps (trx_fuse.h)
static trx_status lowerObjectiveFused(const trxfile::Program &prog, const trxfile::Program &prep,
                                      const trxfile::Program &bprop, std::unique_ptr<trx_program> &fwd,
                                      std::unique_ptr<trx_program> &adj) {
  try {
    trx::FuseOptions opt;
    if (const char *w = getenv("TRX_WARPS")) opt.n_warps = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_RING_SLOTS")) opt.ring_slots = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_SYNC_PENALTY")) opt.sync_penalty = (uint32_t)atoi(w);
    trx::lowerObjective(prog, prep, bprop, opt, fwd->low, adj->low);
  } catch (const std::exception &e) {
    std::string msg = e.what();
    bool unsupported = msg.find("unsupported operator") != std::string::npos;
    return fail(unsupported ? TRX_ERR_UNSUPPORTED : TRX_ERR_INVALID, msg);
  }
  return TRX_OK;
}

#endif

trx_status trx_objective_create(const void *prog_blob, size_t prog_bytes, const void *prep_blob, size_t prep_bytes,
                                const void *bprop_blob, size_t bprop_bytes, uint64_t lanes, uint64_t max_device_bytes,
                                int device, int flags, trx_objective **out) {
  if (!prog_blob || !prep_blob || !bprop_blob || !out) return fail(TRX_ERR_INVALID, "bad argument");
  *out = nullptr;
  if (lanes == 0 || lanes % 8) return fail(TRX_ERR_INVALID, "lanes must be a positive multiple of 8");
  // every early return below releases what has been built so far
  std::unique_ptr<trx_objective, void (*)(trx_objective *)> o(new trx_objective(), trx_objective_destroy);
  o->fused = (flags & TRX_OBJECTIVE_FUSED) != 0;
#ifndef TRX_WITH_EXPERIMENTS
  if (o->fused)
    return fail(TRX_ERR_UNSUPPORTED, "TRX_OBJECTIVE_FUSED: the experimental ring lowering is not built into this library "
                                     "(-DTRX_WITH_EXPERIMENTS)");
#endif
  o->scalar_rows = (flags & TRX_OBJECTIVE_NO_SCALAR_ROWS) == 0;
  {
    trxfile::Program p_prog, p_prep, p_bprop;
    try {
      trxfile::decodeProgram(prog_blob, prog_bytes, p_prog);
      trxfile::decodeProgram(prep_blob, prep_bytes, p_prep);
      trxfile::decodeProgram(bprop_blob, bprop_bytes, p_bprop);
      // the adjoint reads forward operands in place instead of through x_prep's copies (trx_direct.h)
      static int direct_env = getenv("TRX_DIRECT") ? atoi(getenv("TRX_DIRECT")) : 1;
      // (not when a tangent program will be attached: forward_* rules read x_prep's saved operands too)
      if (direct_env && !o->fused && !(flags & TRX_OBJECTIVE_KEEP_PREP_STATE)) trx::directPrepare(p_prep, p_bprop, o->direct);
      o->keeps_prep_state = !(direct_env && !o->fused && !(flags & TRX_OBJECTIVE_KEEP_PREP_STATE));
      static int static_env = getenv("TRX_STATIC_CONSTS") ? atoi(getenv("TRX_STATIC_CONSTS")) : 1;
      o->consts_static = static_env && trx::constantsAreNeverWritten({&p_prog, &p_prep, &p_bprop});
    } catch (const std::exception &e) {
      return fail(TRX_ERR_INVALID, e.what());
    }
    trx_status s;
    trx_program *raw = nullptr;
#ifdef TRX_WITH_EXPERIMENTS
    if (o->fused) {
      std::unique_ptr<trx_program> fwd(new trx_program()), adj(new trx_program());
      if ((s = lowerObjectiveFused(p_prog, p_prep, p_bprop, fwd, adj)) != TRX_OK) return s;
      if ((s = finishProgram(fwd, device, TRX_SCALAR_SHARED, TRX_SCALAR_SHARED, &raw)) != TRX_OK) return s;
      o->prog = raw;
      if ((s = finishProgram(adj, device, TRX_SCALAR_REDUCE, TRX_SCALAR_REDUCE, &raw)) != TRX_OK) return s;
      o->bprop = raw;
      o->prep = nullptr;
    } else
#endif
    if (flags & TRX_OBJECTIVE_SEPARATE) {
      if ((s = createFromFile(p_prog, device, TRX_SCALAR_AUTO, TRX_SCALAR_AUTO, &raw)) != TRX_OK) return s;
      o->prog = raw;
      if ((s = createFromFile(p_prep, device, TRX_SCALAR_AUTO, TRX_SCALAR_AUTO, &raw)) != TRX_OK) return s;
      o->prep = raw;
      if ((s = createFromFile(p_bprop, device, TRX_SCALAR_AUTO, TRX_SCALAR_AUTO, &raw, true)) != TRX_OK) return s;
      o->bprop = raw;
    } else {
      // default: x_prep interleaved into x_prog (every prepare instruction directly after its compute
      // instruction), level-scheduled as one program: the prepare work fills warps that the narrow
      // forward levels leave idle, and one launch disappears
      trxfile::Program combined;
      try {
        trx::combineForward(p_prog, p_prep, combined);
      } catch (const std::exception &e) {
        return fail(TRX_ERR_INVALID, e.what());
      }
      if ((s = createFromFile(combined, device, TRX_SCALAR_SHARED, TRX_SCALAR_SHARED, &raw)) != TRX_OK) return s;
      o->prog = raw;
      o->prep = nullptr;
      if ((s = createFromFile(p_bprop, device, TRX_SCALAR_AUTO, TRX_SCALAR_AUTO, &raw, true)) != TRX_OK) return s;
      o->bprop = raw;
    }
  }
  const trx_program *prog = o->prog, *prep = o->prep, *bprop = o->bprop;
  const auto &po = prog->low.outputs.elems;
  const auto &bi = bprop->low.inputs.elems;
  if (po.size() != bi.size()) return fail(TRX_ERR_INVALID, "bprop inputs do not mirror prog outputs");
  for (auto &e : prog->low.inputs.elems)
    if (e.lane != 0xff) return fail(TRX_ERR_UNSUPPORTED, "objective: lane-typed optimisation variables are not supported");
  if (bprop->low.outputs.elems.size() != prog->low.inputs.elems.size())
    return fail(TRX_ERR_INVALID, "bprop outputs do not mirror prog inputs");
  TRX_CUDA(cudaSetDevice(device));
  o->device = device;
  o->lanes = lanes;
  o->n_x = prog->low.inputs.elems.size();
  // lane chunking: rollouts are independent, so when the write-once memory image of all tiles does
  // not fit the budget the tiles are evaluated in equal chunks and the gradients summed
  uint64_t per_tile = std::max(std::max(prog->low.memory_size, prep ? prep->low.memory_size : 0), bprop->low.memory_size);
  uint64_t tiles = lanes / 8;
  uint64_t budget_tiles = max_device_bytes ? std::max<uint64_t>(1, max_device_bytes / per_tile) : tiles;
  uint64_t chunk_tiles = std::min(tiles, budget_tiles);
  while (tiles % chunk_tiles) chunk_tiles--;
  o->chunk_lanes = chunk_tiles * 8;
  o->n_chunks = tiles / chunk_tiles;
  trx_status s = trx_memory_create(o->chunk_lanes, device, &o->mem);
  if (s != TRX_OK) return s;
  s = ensureTileStride(o->mem, per_tile, 0);
  if (s != TRX_OK) return s;
  std::vector<SeedElem> seed(po.size());
  for (size_t i = 0; i < po.size(); i++) {
    if (po[i].lane != bi[i].lane) return fail(TRX_ERR_INVALID, "bprop inputs do not mirror prog outputs");
    seed[i] = {po[i].idx, bi[i].idx, po[i].lane == 0xff ? 1u : 0u};
  }
  o->n_seed = (uint32_t)seed.size();
  if ((s = upload(seed, &o->d_seed)) != TRX_OK) return s;
  o->param_elems = prog->low.parameters.wideElems(lanes);
  TRX_CUDA(cudaMalloc(&o->d_x, std::max<uint64_t>(1, o->n_x) * 8));
  TRX_CUDA(cudaMalloc(&o->d_params, std::max<uint64_t>(1, o->param_elems) * 8));
  TRX_CUDA(cudaMemset(o->d_params, 0, std::max<uint64_t>(1, o->param_elems) * 8));
  TRX_CUDA(cudaMalloc(&o->d_out, (o->n_x + 1) * 8));
  TRX_CUDA(cudaMalloc(&o->d_partial, tiles * kSeedSplit * 8));
  TRX_CUDA(cudaMallocHost(&o->h_x, std::max<uint64_t>(1, o->n_x) * 8));
  TRX_CUDA(cudaMallocHost(&o->h_out, (o->n_x + 1) * 8));
  for (auto &e : o->ev) TRX_CUDA(cudaEventCreate(&e));
  *out = o.release();
  return TRX_OK;
}

void trx_objective_destroy(trx_objective *o) {
  if (!o) return;
  trxRolloutDestroy(o->rollout);
  trx_program_destroy(o->prog);
  trx_program_destroy(o->prep);
  trx_program_destroy(o->bprop);
  trx_program_destroy(o->fprop);
  cudaFree(o->d_cg);
  cudaFree(o->d_scal);
  if (o->h_scal) cudaFreeHost(o->h_scal);
  trx_memory_destroy(o->mem);
  cudaFree(o->d_x);
  cudaFree(o->d_params);
  cudaFree(o->d_out);
  cudaFree(o->d_partial);
  cudaFree(o->d_seed);
  if (o->h_x) cudaFreeHost(o->h_x);
  if (o->h_out) cudaFreeHost(o->h_out);
  for (auto &e : o->ev)
    if (e) cudaEventDestroy(e);
  for (auto &e : o->prof_events) cudaEventDestroy(e);
  delete o;
}

trx_status trx_objective_info(const trx_objective *o, const char *key, uint64_t *value) {
  if (!o || !key || !value) return fail(TRX_ERR_INVALID, "bad argument");
  std::string k = key;
  if (o->rollout && k != "n_x" && k != "lanes" && k != "parameter_bytes") {
    if (k == "rerolled") { *value = 1; return TRX_OK; }
    uint64_t v = trxRolloutInfo(o->rollout, key);
    if (v == ~0ull) return fail(TRX_ERR_INVALID, "unknown key " + k);
    *value = v;
    return TRX_OK;
  }
  if (k == "rerolled") { *value = 0; return TRX_OK; }
  if (k == "n_x") *value = o->n_x;
  else if (k == "lanes") *value = o->lanes;
  else if (k == "chunk_lanes") *value = o->chunk_lanes;
  else if (k == "n_chunks") *value = o->n_chunks;
  else if (k == "parameter_bytes") *value = o->param_elems * 8;
  else if (k == "device_bytes") *value = o->mem->tile_stride * o->mem->n_tiles * 8;
  else if (k == "fused") *value = o->fused ? 1 : 0;
  else if (k == "consts_static") *value = o->consts_static ? 1 : 0;
  else if (k == "direct.n_prepare_dropped") *value = o->direct.n_prepare_dropped;
  else if (k == "direct.n_arguments_redirected") *value = o->direct.n_arguments_redirected;
  else if (k == "direct.n_mul_rewritten") *value = o->direct.n_mul_rewritten;
  else if (k.rfind("prog.", 0) == 0) return trx_program_stat(o->prog, key + 5, value);
  else if (k.rfind("prep.", 0) == 0) {
    if (!o->prep) { *value = 0; return TRX_OK; }
    return trx_program_stat(o->prep, key + 5, value);
  }
  else if (k.rfind("bprop.", 0) == 0) return trx_program_stat(o->bprop, key + 6, value);
  else if (k == "launches_per_evaluation") {
    uint64_t per_chunk = 0;
    per_chunk += o->param_elems ? 1 : 0;                       // parameter scatter
    per_chunk += 1;                                            // input scatter
    per_chunk += (o->prog->low.const_idx.empty() ? 0 : 1) + 1; // prog
    per_chunk += 1;                                            // seed
    per_chunk += (o->prep && o->prep->low.n_instructions ? 1 : 0);  // prep
    per_chunk += (o->bprop->low.n_instructions ? 1 : 0);       // bprop
    per_chunk += 1;                                            // gradient gather
    *value = per_chunk * o->n_chunks + 1;
  } else return fail(TRX_ERR_INVALID, "unknown key " + k);
  return TRX_OK;
}

// per-lane parameters of all `lanes` rollouts (wide layout), host pointer
trx_status trx_objective_parameterize(trx_objective *o, const void *packed, uint64_t bytes, void *stream_) {
  if (!o || (!packed && bytes)) return fail(TRX_ERR_INVALID, "bad argument");
  if (o->rollout) {
    std::string err;
    trx_status s = trxRolloutParameterize(o->rollout, packed, bytes, (cudaStream_t)stream_, err);
    return s == TRX_OK ? s : fail(s, err);
  }
  if (bytes != o->param_elems * 8)
    return fail(TRX_ERR_SIZE, "packed buffer size mismatch: got " + std::to_string(bytes) + " expected " +
                                  std::to_string(o->param_elems * 8));
  TRX_CUDA(cudaSetDevice(o->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (bytes) {
    TRX_CUDA(cudaMemcpyAsync(o->d_params, packed, bytes, cudaMemcpyHostToDevice, stream));
    TRX_CUDA(cudaStreamSynchronize(stream));
  }
  return TRX_OK;
}

// The hot path.  x on the device (n_x doubles) -> d_out = [loss, g_0 .. g_{n_x-1}] on the device.
// Sequence per chunk = GradientDescentSolver::_step's engine calls (solvers/gd.h:50-63):
// x_prog input+execute, loss = |r|^2, x_prep execute, x_bprop input(r)+execute+output.
// seed of the in-tape random operators; the execute counter restarts, so a run is reproducible from here
void trx_set_random_seed(uint64_t seed) {
  g_rng_seed.store(seed);
  g_rng_epoch.store(0);
}

// ---- collision shapes
trx_status trx_collision_shape_create(const trx_shape_desc *d, uint64_t *handle) {
  if (!d || !handle) return fail(TRX_ERR_INVALID, "bad argument");
  ShapeRegistry &R = shapeRegistry();
  std::lock_guard<std::mutex> lock(R.mu);
  bool unsupported = false;
  const std::string err = R.table.addShape(*d, *handle, unsupported);
  if (!err.empty()) return fail(unsupported ? TRX_ERR_UNSUPPORTED : TRX_ERR_INVALID, err);
  R.version++;
  trxvm::shapeHeapHost() = R.table.heap.data();
  return TRX_OK;
}

trx_status trx_collision_pair_create(uint64_t a, uint64_t b, uint64_t *handle) {
  if (!handle) return fail(TRX_ERR_INVALID, "bad argument");
  ShapeRegistry &R = shapeRegistry();
  std::lock_guard<std::mutex> lock(R.mu);
  const std::string err = R.table.addPair(a, b, *handle);
  if (!err.empty()) return fail(TRX_ERR_INVALID, err);
  R.version++;
  trxvm::shapeHeapHost() = R.table.heap.data();
  return TRX_OK;
}

trx_status trx_collision_shape_center(uint64_t shape, double *center3) {
  ShapeRegistry &R = shapeRegistry();
  std::lock_guard<std::mutex> lock(R.mu);
  if (!center3 || std::find(R.table.shapes.begin(), R.table.shapes.end(), shape) == R.table.shapes.end())
    return fail(TRX_ERR_INVALID, "not a shape handle");
  for (int k = 0; k < 3; k++) center3[k] = R.table.heap[shape + trxcol::REC_CENTER + k];
  return TRX_OK;
}

void trx_collision_clear(void) {
  ShapeRegistry &R = shapeRegistry();
  std::lock_guard<std::mutex> lock(R.mu);
  R.table.clear();
  R.version++;
  R.generation++;
  trxvm::shapeHeapHost() = R.table.heap.data();
}

trx_status trx_collision_table(double *out, uint64_t capacity, uint64_t *count) {
  if (!count) return fail(TRX_ERR_INVALID, "bad argument");
  ShapeRegistry &R = shapeRegistry();
  std::lock_guard<std::mutex> lock(R.mu);
  *count = R.table.heap.size();
  if (out) std::memcpy(out, R.table.heap.data(), std::min<uint64_t>(capacity, *count) * sizeof(double));
  return TRX_OK;
}

__global__ void trx_collision_pairs_kernel(uint64_t pair, const double *poses, uint64_t n, double *out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *q = poses + 14 * i;
  trxvm::PoseV a, b;
  a.t = trxvm::V3{q[0], q[1], q[2]}, a.r = trxvm::Q4{q[3], q[4], q[5], q[6]};
  b.t = trxvm::V3{q[7], q[8], q[9]}, b.r = trxvm::Q4{q[10], q[11], q[12], q[13]};
  trxcol::Result r;
  trxcol::collideShapes(trxvm::g_shape_heap_dev, pair, a, b, r);
  double *o = out + 10 * i;
  o[0] = r.a.x, o[1] = r.a.y, o[2] = r.a.z, o[3] = r.b.x, o[4] = r.b.y, o[5] = r.b.z;
  o[6] = r.n.x, o[7] = r.n.y, o[8] = r.n.z, o[9] = r.d;
}
__global__ void trx_collision_project_kernel(uint64_t shape, const double *points, uint64_t n, double *out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  trxvm::V3 nrm = trxvm::V3{0.0, 0.0, 0.0};
  double d = 0.0;
  trxcol::projectPoint(trxvm::g_shape_heap_dev, shape, trxvm::V3{points[3 * i], points[3 * i + 1], points[3 * i + 2]}, nrm, d);
  out[4 * i] = nrm.x, out[4 * i + 1] = nrm.y, out[4 * i + 2] = nrm.z, out[4 * i + 3] = d;
}

static trx_status collisionBatch(bool pairs, uint64_t handle, const double *in, uint64_t n, double *out, int device) {
  if ((!in || !out) && n) return fail(TRX_ERR_INVALID, "bad argument");
  {
    ShapeRegistry &R = shapeRegistry();
    std::lock_guard<std::mutex> lock(R.mu);
    const auto &known = pairs ? R.table.pairs : R.table.shapes;
    if (std::find(known.begin(), known.end(), handle) == known.end()) return fail(TRX_ERR_INVALID, "unknown handle");
  }
  if (n == 0) return TRX_OK;
  TRX_CUDA(cudaSetDevice(device));
  trx_status s = syncShapeTable(device, 0);
  if (s != TRX_OK) return s;
  const size_t in_w = pairs ? 14 : 3, out_w = pairs ? 10 : 4;
  double *d_in = nullptr, *d_out = nullptr;
  TRX_CUDA(cudaMalloc(&d_in, n * in_w * sizeof(double)));
  if (cudaMalloc(&d_out, n * out_w * sizeof(double)) != cudaSuccess) {
    cudaFree(d_in);
    return fail(TRX_ERR_CUDA, "out of device memory");
  }
  cudaMemcpy(d_in, in, n * in_w * sizeof(double), cudaMemcpyHostToDevice);
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (pairs) trx_collision_pairs_kernel<<<blocks, 128>>>(handle, d_in, n, d_out);
  else trx_collision_project_kernel<<<blocks, 128>>>(handle, d_in, n, d_out);
  g_launches++;
  cudaError_t e = cudaMemcpy(out, d_out, n * out_w * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(TRX_ERR_CUDA, cudaGetErrorString(e));
  return TRX_OK;
}
trx_status trx_collision_query_pairs(uint64_t pair, const double *poses, uint64_t n, double *out, int device) {
  return collisionBatch(true, pair, poses, n, out, device);
}
trx_status trx_collision_project_points(uint64_t shape, const double *points, uint64_t n, double *out, int device) {
  return collisionBatch(false, shape, points, n, out, device);
}

// ---- communicator
trx_status trx_comm_unique_id(uint8_t *id) {
  if (!id) return fail(TRX_ERR_INVALID, "bad argument");
  const trxcomm::Api &nccl = trxcomm::api();
  if (!nccl.error.empty()) return fail(TRX_ERR_UNSUPPORTED, nccl.error);
  trxcomm::UniqueId uid;
  int rc = nccl.GetUniqueId(&uid);
  if (rc != trxcomm::kNcclSuccess) return fail(TRX_ERR_CUDA, std::string("ncclGetUniqueId: ") + nccl.GetErrorString(rc));
  std::memcpy(id, uid.internal, TRX_COMM_ID_BYTES);
  return TRX_OK;
}
trx_status trx_comm_create(const uint8_t *id, int n_ranks, int rank, int device, trx_comm **out) {
  if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(TRX_ERR_INVALID, "bad argument");
  *out = nullptr;
  const trxcomm::Api &nccl = trxcomm::api();
  if (!nccl.error.empty()) return fail(TRX_ERR_UNSUPPORTED, nccl.error);
  TRX_CUDA(cudaSetDevice(device));
  trxcomm::UniqueId uid;
  std::memcpy(uid.internal, id, TRX_COMM_ID_BYTES);
  std::unique_ptr<trx_comm> c(new trx_comm());
  int rc = nccl.CommInitRank(&c->comm, n_ranks, uid, rank);
  if (rc != trxcomm::kNcclSuccess) return fail(TRX_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl.GetErrorString(rc));
  c->n_ranks = n_ranks, c->rank = rank, c->device = device;
  *out = c.release();
  return TRX_OK;
}
void trx_comm_destroy(trx_comm *c) {
  if (!c) return;
  if (c->comm) trxcomm::api().CommDestroy(c->comm);
  delete c;
}
trx_status trx_comm_allreduce(trx_comm *c, double *d_values, uint64_t count, void *stream_) {
  if (!c || !d_values) return fail(TRX_ERR_INVALID, "bad argument");
  const trxcomm::Api &nccl = trxcomm::api();
  int rc = nccl.AllReduce(d_values, d_values, count, trxcomm::kNcclFloat64, trxcomm::kNcclSum, c->comm, (cudaStream_t)stream_);
  if (rc != trxcomm::kNcclSuccess) return fail(TRX_ERR_CUDA, std::string("ncclAllReduce: ") + nccl.GetErrorString(rc));
  return TRX_OK;
}
trx_status trx_objective_set_comm(trx_objective *o, trx_comm *comm) {
  if (!o) return fail(TRX_ERR_INVALID, "bad argument");
  if (comm && comm->device != o->device) return fail(TRX_ERR_INVALID, "communicator and objective live on different devices");
  o->comm = comm;
  return TRX_OK;
}

// this rank's share: d_out = {loss, gradient} over the rollouts of this objective
static trx_status evaluateLocal(trx_objective *o, const double *d_x, double *d_out, void *stream_);

trx_status trx_objective_evaluate_device(trx_objective *o, const double *d_x, double *d_out, void *stream_) {
  if (!o || !d_x || !d_out) return fail(TRX_ERR_INVALID, "bad argument");
  trx_status s = evaluateLocal(o, d_x, d_out, stream_);
  if (s != TRX_OK || !o->comm) return s;
  // rollouts are sharded over the ranks: one sum of {loss, gradient}, on the same stream right behind the kernels
  return trx_comm_allreduce(o->comm, d_out, o->n_x + 1, stream_);
}

static trx_status evaluateLocal(trx_objective *o, const double *d_x, double *d_out, void *stream_) {
  TRX_CUDA(cudaSetDevice(o->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (o->rollout) {
    std::string err;
    uint64_t before = trxRolloutLaunches(o->rollout);
    trx_status s = trxRolloutEvaluate(o->rollout, d_x, d_out, nullptr, nullptr, true, stream, err);
    g_launches += trxRolloutLaunches(o->rollout) - before;
    return s == TRX_OK ? s : fail(s, err);
  }
  trx_memory *m = o->mem;
  const uint64_t chunk_tiles = m->n_tiles;
  for (uint64_t c = 0; c < o->n_chunks; c++) {
    trx_status s;
    if (o->param_elems) {
      s = scatterDevice(o->prog, m, TRX_BUF_PARAMETER, o->d_params, o->param_elems * 8, stream, o->lanes,
                        c * o->chunk_lanes);
      if (s != TRX_OK) return s;
    }
    s = scatterDevice(o->prog, m, TRX_BUF_INPUT, d_x, o->n_x * 8, stream, o->lanes, c * o->chunk_lanes);
    if (s != TRX_OK) return s;
    m->rng_lane_base = c * o->chunk_lanes;
    cudaEvent_t *pe = o->profiling ? &o->prof_events[c * 4] : nullptr;
    if (pe) cudaEventRecord(pe[0], stream);
    // constants: uploaded by every execute as the reference does (simple.cpp:27-30), except when this
    // objective has proven that nothing it runs can overwrite one -- then once is enough
    const bool consts = !o->consts_resident;
    if ((s = executeImpl(o->prog, m, stream_, consts)) != TRX_OK) return s;
    if (pe) cudaEventRecord(pe[1], stream);
    trx_seed_kernel<<<dim3((unsigned)chunk_tiles, kSeedSplit), 256, 0, stream>>>(
        o->d_seed, o->n_seed, m->d_mem, m->tile_stride, (c == 0 && o->scalar_rows) ? 1 : 0,
        o->d_partial + c * chunk_tiles * kSeedSplit);
    g_launches++;
    if (o->prep && (s = executeImpl(o->prep, m, stream_, consts)) != TRX_OK) return s;
    if (pe) cudaEventRecord(pe[2], stream);
    if ((s = executeImpl(o->bprop, m, stream_, consts)) != TRX_OK) return s;
    if (pe) cudaEventRecord(pe[3], stream);
    if (c + 1 == o->n_chunks && o->consts_static) o->consts_resident = true;
    uint32_t n = o->bprop->n_elems[TRX_BUF_OUTPUT];
    dim3 grid((n + 255) / 256, 1);
    trx_gather_kernel<<<grid, 256, 0, stream>>>(o->bprop->d_elems[TRX_BUF_OUTPUT], n, d_out + 1, o->lanes, 0, m->d_mem,
                                               m->tile_stride, (uint32_t)chunk_tiles, 1, c == 0 ? 0 : 1);
    g_launches++;
  }
  trx_loss_kernel<<<1, 256, 0, stream>>>(o->d_partial, (uint32_t)(o->lanes / 8) * kSeedSplit, d_out);
  g_launches++;
  TRX_CUDA(cudaGetLastError());
  return TRX_OK;
}

// Host-buffer variant (the reference-facing call): copies x from `x` (n_x doubles) through pinned
// memory, evaluates, and returns loss and gradient (n_x doubles) to the host; synchronous.
trx_status trx_objective_evaluate(trx_objective *o, const double *x, double *loss, double *gradient, void *stream_) {
  if (!o || !x || !loss || !gradient) return fail(TRX_ERR_INVALID, "bad argument");
  TRX_CUDA(cudaSetDevice(o->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  std::memcpy(o->h_x, x, o->n_x * 8);
  TRX_CUDA(cudaEventRecord(o->ev[0], stream));
  TRX_CUDA(cudaMemcpyAsync(o->d_x, o->h_x, o->n_x * 8, cudaMemcpyHostToDevice, stream));
  TRX_CUDA(cudaEventRecord(o->ev[1], stream));
  trx_status s = trx_objective_evaluate_device(o, o->d_x, o->d_out, stream_);
  if (s != TRX_OK) return s;
  TRX_CUDA(cudaEventRecord(o->ev[2], stream));
  TRX_CUDA(cudaMemcpyAsync(o->h_out, o->d_out, (o->n_x + 1) * 8, cudaMemcpyDeviceToHost, stream));
  TRX_CUDA(cudaEventRecord(o->ev[3], stream));
  TRX_CUDA(cudaStreamSynchronize(stream));
  *loss = o->h_out[0];
  std::memcpy(gradient, o->h_out + 1, o->n_x * 8);
  cudaEventElapsedTime(&o->last_ms[0], o->ev[0], o->ev[1]);
  cudaEventElapsedTime(&o->last_ms[1], o->ev[1], o->ev[2]);
  cudaEventElapsedTime(&o->last_ms[2], o->ev[2], o->ev[3]);
  return TRX_OK;
}

// Per-program kernel timing.  enable: CUDA events are recorded around the prog / prep(+seed) / bprop
// launches of every chunk; trx_objective_profile_collect (after a stream synchronise) adds the
// elapsed times of the last evaluation to running totals {prog, prep, bprop} in milliseconds.
trx_status trx_objective_set_profiling(trx_objective *o, int enable) {
  if (!o) return fail(TRX_ERR_INVALID, "bad argument");
  if (o->rollout) return fail(TRX_ERR_UNSUPPORTED, "the re-rolled objective is one launch: time it with events on the stream");
  TRX_CUDA(cudaSetDevice(o->device));
  o->profiling = enable != 0;
  if (o->profiling && o->prof_events.empty()) {
    o->prof_events.resize(o->n_chunks * 4);
    for (auto &e : o->prof_events) TRX_CUDA(cudaEventCreate(&e));
  }
  o->prof_ms[0] = o->prof_ms[1] = o->prof_ms[2] = 0;
  o->prof_count = 0;
  return TRX_OK;
}
trx_status trx_objective_profile_collect(trx_objective *o, double *ms3, uint64_t *evaluations) {
  if (!o || !ms3) return fail(TRX_ERR_INVALID, "bad argument");
  if (!o->profiling) return fail(TRX_ERR_INVALID, "profiling is not enabled");
  TRX_CUDA(cudaSetDevice(o->device));
  for (uint64_t c = 0; c < o->n_chunks; c++) {
    cudaEvent_t *pe = &o->prof_events[c * 4];
    TRX_CUDA(cudaEventSynchronize(pe[3]));
    for (int k = 0; k < 3; k++) {
      float ms = 0;
      TRX_CUDA(cudaEventElapsedTime(&ms, pe[k], pe[k + 1]));
      o->prof_ms[k] += ms;
    }
  }
  o->prof_count++;
  for (int k = 0; k < 3; k++) ms3[k] = o->prof_ms[k];
  if (evaluations) *evaluations = o->prof_count;
  return TRX_OK;
}

// n_steps gradient-descent steps entirely on the device: no host round trip between the steps.
// d_x[n_x] (in/out) optimisation variables, d_v[n_x] (in/out) momentum state, d_losses[n_steps] (out, may be
// null) the loss BEFORE each update, as GradientDescentSolver::loss() reports it.  Asynchronous on `stream`.
trx_status trx_objective_gd_steps(trx_objective *o, double *d_x, double *d_v, double learning_rate, double momentum,
                                  uint32_t n_steps, double *d_losses, void *stream_) {
  if (!o || !d_x || !d_v) return fail(TRX_ERR_INVALID, "bad argument");
  TRX_CUDA(cudaSetDevice(o->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  for (uint32_t it = 0; it < n_steps; it++) {
    trx_status s = trx_objective_evaluate_device(o, d_x, o->d_out, stream_);
    if (s != TRX_OK) return s;
    trx_gd_update_kernel<<<1, 1024, 0, stream>>>(d_x, d_v, o->d_out, (uint32_t)o->n_x, learning_rate, momentum,
                                                d_losses ? d_losses + it : nullptr);
    g_launches++;
  }
  TRX_CUDA(cudaGetLastError());
  return TRX_OK;
}

// ---- Gauss-Newton: tangent sweep, J^T J v, LeastSquaresSolver::_step
trx_status trx_objective_attach_fprop(trx_objective *o, const void *fprop_blob, size_t fprop_bytes) {
  if (!o || !fprop_blob) return fail(TRX_ERR_INVALID, "bad argument");
  if (o->rollout)
    return fail(TRX_ERR_UNSUPPORTED, "the re-rolled objective has no tangent sweep: Gauss-Newton products need the tape objective");
  if (o->fused) return fail(TRX_ERR_UNSUPPORTED, "fused lowering keeps no tangent slots");
  if (!o->keeps_prep_state)
    return fail(TRX_ERR_INVALID,
                "this objective dropped x_prep's operand copies (the adjoint reads forward values in place); create it "
                "with TRX_OBJECTIVE_KEEP_PREP_STATE: the tangent rules read the saved operands");
  trxfile::Program p_fprop;
  try {
    trxfile::decodeProgram(fprop_blob, fprop_bytes, p_fprop);
  } catch (const std::exception &e) {
    return fail(TRX_ERR_INVALID, e.what());
  }
  trx_program *raw = nullptr;
  trx_status s = createFromFile(p_fprop, o->device, TRX_SCALAR_SHARED, TRX_SCALAR_SHARED, &raw);
  if (s != TRX_OK) return s;
  if (raw->low.inputs.elems.size() != o->prog->low.inputs.elems.size() ||
      raw->low.outputs.elems.size() != o->bprop->low.inputs.elems.size()) {
    trx_program_destroy(raw);
    return fail(TRX_ERR_INVALID, "fprop ports do not mirror the objective's");
  }
  for (size_t i = 0; i < raw->low.outputs.elems.size(); i++)
    if (raw->low.outputs.elems[i].idx != o->bprop->low.inputs.elems[i].idx) {
      // hprop = fprop ; bprop works because the tangent of an output lands in the adjoint's seed slot (gradients.cpp:344-362)
      trx_program_destroy(raw);
      return fail(TRX_ERR_INVALID, "fprop outputs are not the adjoint's seed slots");
    }
  trx_program_destroy(o->fprop);
  o->fprop = raw;
  // the tangent sweep writes slots of its own; the objective's constants are uploaded by every execute again
  o->consts_static = false;
  o->consts_resident = false;
  TRX_CUDA(cudaSetDevice(o->device));
  return ensureTileStride(o->mem, raw->low.memory_size, nullptr);
}

// d_hv[n_x] = J^T J d_v at the point of the last evaluation (x_prog and x_prep state are in the memory image):
// x_fprop->input(v); execute; x_bprop->execute; output -- the reference's x_hprop (core/eigen.h:73-77), with the
// lane-independent rows seeded once and the sum over tiles (and ranks) that completes reverse_batch.
trx_status trx_objective_hessian_product(trx_objective *o, const double *d_v, double *d_hv, void *stream_) {
  if (!o || !d_v || !d_hv) return fail(TRX_ERR_INVALID, "bad argument");
  if (!o->fprop) return fail(TRX_ERR_INVALID, "no tangent program: call trx_objective_attach_fprop first");
  if (o->n_chunks != 1)
    return fail(TRX_ERR_UNSUPPORTED, "Gauss-Newton products need all rollouts resident (one chunk): raise max_device_bytes");
  TRX_CUDA(cudaSetDevice(o->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  trx_memory *m = o->mem;
  trx_status s = scatterDevice(o->fprop, m, TRX_BUF_INPUT, d_v, o->n_x * 8, stream, o->lanes, 0);
  if (s != TRX_OK) return s;
  if ((s = executeImpl(o->fprop, m, stream_, true)) != TRX_OK) return s;
  {
    const uint32_t n = o->bprop->n_elems[TRX_BUF_INPUT];
    trx_mask_scalar_kernel<<<dim3((n + 255) / 256, (unsigned)m->n_tiles), 256, 0, stream>>>(
        o->bprop->d_elems[TRX_BUF_INPUT], n, m->d_mem, m->tile_stride, o->scalar_rows ? 1 : 0);
    g_launches++;
  }
  if ((s = executeImpl(o->bprop, m, stream_, true)) != TRX_OK) return s;
  {
    const uint32_t n = o->bprop->n_elems[TRX_BUF_OUTPUT];
    trx_gather_kernel<<<dim3((n + 255) / 256, 1), 256, 0, stream>>>(o->bprop->d_elems[TRX_BUF_OUTPUT], n, d_hv, o->lanes, 0,
                                                                   m->d_mem, m->tile_stride, (uint32_t)m->n_tiles, 1, 0);
    g_launches++;
  }
  TRX_CUDA(cudaGetLastError());
  if (o->comm) return trx_comm_allreduce(o->comm, d_hv, o->n_x, stream_);
  return TRX_OK;
}

// LeastSquaresSolver::_step (solvers/sq.h:57-123) n_steps times with the vectors resident on the device:
//   r = x_prog(x), loss = |r|^2, x_prep, g = x_bprop(r)                       (trx_objective_evaluate_device)
//   solve (J^T J + lambda I) d = g by conjugate gradients                      (Eigen::ConjugateGradient, identity
//     preconditioner, zero initial guess, stop when |residual|^2 < tol^2 |g|^2 or after max_linear_iterations;
//     product = trx_objective_hessian_product + lambda v, solvers/base.h:76-85)
//   x <- x (+) (-step_scaling d)                                               (sq.h:101,121-122; scalar inputs: add)
// The host sees two dot products per linear iteration and nothing else.  losses[n_steps] (host, may be null): the
// loss before each update; linear_iterations[n_steps] (host, may be null).  Synchronous.
trx_status trx_objective_sq_steps(trx_objective *o, double *d_x, double regularization, uint32_t max_linear_iterations,
                                  double step_scaling, double tolerance, uint32_t n_steps, double *losses,
                                  uint32_t *linear_iterations, void *stream_) {
  if (!o || !d_x) return fail(TRX_ERR_INVALID, "bad argument");
  if (!o->fprop) return fail(TRX_ERR_INVALID, "no tangent program: call trx_objective_attach_fprop first");
  TRX_CUDA(cudaSetDevice(o->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const uint32_t n = (uint32_t)o->n_x;
  if (!o->d_cg) {
    TRX_CUDA(cudaMalloc(&o->d_cg, (size_t)5 * n * sizeof(double)));
    TRX_CUDA(cudaMalloc(&o->d_scal, 8 * sizeof(double)));
    TRX_CUDA(cudaMallocHost(&o->h_scal, 8 * sizeof(double)));
  }
  double *rhs = o->d_cg, *sol = rhs + n, *res = sol + n, *dir = res + n, *tmp = dir + n;
  const unsigned blocks = (n + 255) / 256;
  auto dot = [&](const double *a, const double *b, double &value) -> trx_status {
    trx_dot_kernel<<<1, 1024, 0, stream>>>(a, b, n, o->d_scal);
    g_launches++;
    TRX_CUDA(cudaMemcpyAsync(o->h_scal, o->d_scal, sizeof(double), cudaMemcpyDeviceToHost, stream));
    TRX_CUDA(cudaStreamSynchronize(stream));
    value = o->h_scal[0];
    return TRX_OK;
  };
  const uint32_t max_it = max_linear_iterations ? max_linear_iterations : n;  // sq.h:86-90
  for (uint32_t step = 0; step < n_steps; step++) {
    trx_status s = trx_objective_evaluate_device(o, d_x, o->d_out, stream_);
    if (s != TRX_OK) return s;
    TRX_CUDA(cudaMemcpyAsync(rhs, o->d_out + 1, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    TRX_CUDA(cudaMemcpyAsync(res, rhs, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    TRX_CUDA(cudaMemsetAsync(sol, 0, n * sizeof(double), stream));
    TRX_CUDA(cudaMemcpyAsync(o->h_scal + 1, o->d_out, sizeof(double), cudaMemcpyDeviceToHost, stream));
    double rhs_norm2 = 0;
    if ((s = dot(rhs, rhs, rhs_norm2)) != TRX_OK) return s;
    if (losses) losses[step] = o->h_scal[1];
    if (!std::isfinite(rhs_norm2)) return fail(TRX_ERR_INVALID, "gradient not finite (sq.h:76)");
    uint32_t it = 0;
    if (rhs_norm2 > 0) {
      const double threshold = std::max(tolerance * tolerance * rhs_norm2, std::numeric_limits<double>::min());
      double residual_norm2 = rhs_norm2;
      if (residual_norm2 >= threshold) {
        TRX_CUDA(cudaMemcpyAsync(dir, res, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        double abs_new = residual_norm2;
        while (it < max_it) {
          if ((s = trx_objective_hessian_product(o, dir, tmp, stream_)) != TRX_OK) return s;
          if (regularization > 0) {
            trx_axpy2_kernel<<<blocks, 256, 0, stream>>>(tmp, regularization, dir, nullptr, 0.0, n);
            g_launches++;
          }
          double p_dot_tmp = 0;
          if ((s = dot(dir, tmp, p_dot_tmp)) != TRX_OK) return s;
          const double alpha = abs_new / p_dot_tmp;
          trx_axpy2_kernel<<<blocks, 256, 0, stream>>>(sol, alpha, dir, nullptr, 0.0, n);
          trx_axpy2_kernel<<<blocks, 256, 0, stream>>>(res, -alpha, tmp, nullptr, 0.0, n);
          g_launches += 2;
          if ((s = dot(res, res, residual_norm2)) != TRX_OK) return s;
          if (residual_norm2 < threshold) break;
          const double abs_old = abs_new;
          abs_new = residual_norm2;
          trx_cg_direction_kernel<<<blocks, 256, 0, stream>>>(dir, res, abs_new / abs_old, n);
          g_launches++;
          it++;
        }
      }
    }
    if (linear_iterations) linear_iterations[step] = it;
    trx_axpy2_kernel<<<blocks, 256, 0, stream>>>(d_x, -step_scaling, sol, nullptr, 0.0, n);
    g_launches++;
  }
  TRX_CUDA(cudaGetLastError());
  TRX_CUDA(cudaStreamSynchronize(stream));
  return TRX_OK;
}

// The re-rolled form of the objective (csrc/rollout, trx_rollout.cu): built from the problem description
// instead of from flat tapes; every other trx_objective_* call works on it unchanged.
trx_status trx_objective_create_rollout(const char *options, uint64_t lanes, uint64_t max_device_bytes, int device,
                                        int flags, trx_objective **out) {
  if (!options || !out) return fail(TRX_ERR_INVALID, "bad argument");
  *out = nullptr;
  std::unique_ptr<trx_objective, void (*)(trx_objective *)> o(new trx_objective(), trx_objective_destroy);
  o->scalar_rows = (flags & TRX_OBJECTIVE_NO_SCALAR_ROWS) == 0;
  o->device = device;
  o->lanes = lanes;
  std::string err;
  const int checkpoint_mode = (flags & TRX_OBJECTIVE_CHECKPOINT_STATE) ? 0 : (flags & TRX_OBJECTIVE_CHECKPOINT_POLICY) ? 1 : 2;
  trx_status s = trxRolloutCreate(options, lanes, max_device_bytes, device, o->scalar_rows, checkpoint_mode, &o->rollout, err);
  if (s != TRX_OK) return fail(s, err);
  o->n_x = trxRolloutInfo(o->rollout, "n_x");
  o->param_elems = trxRolloutInfo(o->rollout, "parameter_elems");
  o->chunk_lanes = trxRolloutInfo(o->rollout, "chunk_lanes");
  o->n_chunks = trxRolloutInfo(o->rollout, "n_chunks");
  TRX_CUDA(cudaMalloc(&o->d_x, std::max<uint64_t>(1, o->n_x) * 8));
  TRX_CUDA(cudaMalloc(&o->d_out, (o->n_x + 1) * 8));
  TRX_CUDA(cudaMallocHost(&o->h_x, std::max<uint64_t>(1, o->n_x) * 8));
  TRX_CUDA(cudaMallocHost(&o->h_out, (o->n_x + 1) * 8));
  for (auto &e : o->ev) TRX_CUDA(cudaEventCreate(&e));
  *out = o.release();
  return TRX_OK;
}

// r = x_prog->run(x) (solvers/gd.h:50-53) for all rollouts, to the host; wide layout
trx_status trx_objective_residuals(trx_objective *o, const double *x, double *r, uint64_t r_bytes, void *stream_) {
  if (!o || !x || !r) return fail(TRX_ERR_INVALID, "bad argument");
  if (!o->rollout) return fail(TRX_ERR_UNSUPPORTED, "tape objectives: run the prog executable (trx_execute / trx_output)");
  TRX_CUDA(cudaSetDevice(o->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const uint64_t n = trxRolloutInfo(o->rollout, "residual_elems");
  if (r_bytes != n * 8)
    return fail(TRX_ERR_SIZE, "residual buffer size mismatch: got " + std::to_string(r_bytes) + " expected " + std::to_string(n * 8));
  double *d_r = nullptr;
  TRX_CUDA(cudaMalloc(&d_r, n * 8));
  std::memcpy(o->h_x, x, o->n_x * 8);
  cudaError_t e = cudaMemcpyAsync(o->d_x, o->h_x, o->n_x * 8, cudaMemcpyHostToDevice, stream);
  std::string err;
  trx_status s = TRX_OK;
  if (e == cudaSuccess) {
    uint64_t before = trxRolloutLaunches(o->rollout);
    s = trxRolloutEvaluate(o->rollout, o->d_x, o->d_out, d_r, nullptr, false, stream, err);
    g_launches += trxRolloutLaunches(o->rollout) - before;
  }
  if (e == cudaSuccess && s == TRX_OK) e = cudaMemcpyAsync(r, d_r, n * 8, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess && s == TRX_OK) e = cudaStreamSynchronize(stream);
  cudaFree(d_r);
  if (s != TRX_OK) return fail(s, err);
  if (e != cudaSuccess) return fail(TRX_ERR_CUDA, cudaGetErrorString(e));
  return TRX_OK;
}

// device times of the last trx_objective_evaluate call: h2d, kernels, d2h (milliseconds, CUDA events)
trx_status trx_objective_last_times(const trx_objective *o, float *ms3) {
  if (!o || !ms3) return fail(TRX_ERR_INVALID, "bad argument");
  ms3[0] = o->last_ms[0];
  ms3[1] = o->last_ms[1];
  ms3[2] = o->last_ms[2];
  return TRX_OK;
}

}  // extern "C"
