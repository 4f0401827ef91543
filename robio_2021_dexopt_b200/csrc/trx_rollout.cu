// trx_rollout.cu -- the re-rolled rollout objective on sm_100a: one kernel evaluates residuals, loss and
// J^T r of dexopt's policy-rollout objective for all rollouts, frame by frame, with every per-frame
// intermediate on chip.
//
//   reference                                          here
//   -------------------------------------------------  -------------------------------------------------
//   x_prog / x_prep / x_bprop executables over a flat  one launch of trx_rollout_kernel: forward sweep over the
//   write-once tape image (solvers/gd.h:50-63,         frames (state checkpoint per frame -> HBM), then the
//   src/engines/simple.cpp:22-35)                      adjoint sweep, which recomputes each frame from its
//                                                      checkpoint (csrc/rollout/rollout_frame.h)
//   batch + dot4add + madd per weight                  policy layers as FP64 tensor-core tiles (DMMA m8n8k4) over
//   (dexopt/src/tensor.h:134-172, ops.h:22-85)         the rollouts of the CTA, weights resident in shared memory
//   reverse_batch lane sums (dexopt/src/ops.h:22-34)   dW = X^T G accumulated in registers over all frames and
//                                                      rollouts of the CTA, one deterministic cross-CTA sum at the end
//
// Mapping: a CTA works on a group of NW = 8 rollouts with 8 main warps and 8 helper warps (kernel comment below).
// Main warps run the frame phases with thread = (rollout, item slot) -- the lanes of a warp hold the same item of
// the 8 rollouts -- the forward kinematics and its adjoint as warp scans (warp = rollout), and the policy layers as
// tile workers of one [8 x n_in] x [n_in x n_out] product (M = the rollouts).  Helper warps take what is off the
// frame's critical path: the body step and its adjoint, and the weight gradients.  HBM sees the checkpoint
// (written once, read once), the weights (once per CTA) and the per-CTA partial gradient.
//
// sm_100a only; no CPU path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/trx_b200.h"
#include "assembly/dexsyn_options.h"
#include "rollout/rollout_frame.h"
#include "trx_rollout.h"

using namespace rollout;

namespace {

struct KHeader {
  int32_t v[H_COUNT];
};

struct KArgs {
  const int32_t *I;  // plan integer tables (header included), ni words
  const double *D;   // plan double tables, nd doubles
  const double *x;   // policy parameters
  const double *params;  // [3][R] per-rollout parameters (object start offset x, y, yaw)
  double *ckpt;      // [chunk rollouts (padded)][frames][ckpt doubles]
  double *partial;   // [grid][1 + nx]: loss and gradient partial of each CTA
  double *r_out;     // residual vector (wide layout: scalar rows, then [row][R]) or null
  const double *seed;  // adjoint seeds in the same layout, or null (seeds are the residuals)
  long R;            // rollouts of the whole objective (stride of r_out / params)
  long first;        // first rollout of this launch (chunk)
  long count;        // rollouts of this launch
  int ni, nd, nx;
  int frames, n_groups, want_grad, accumulate;
  int ckpt_doubles;  // doubles per (rollout, frame) checkpoint
  int ckpt_mode;     // what a checkpoint holds beyond the carried state: 0 nothing, 1 the policy values, 2 everything the adjoint reads
  // shared-memory layout, in doubles
  int o_D, o_I, o_W1, o_B1, o_W2, o_B2, s_W1, s_W2, r_W1, r_W2;
  int o_X, o_H, o_A, o_O, o_GX, o_GH, o_GO, s_X, s_H, s_O;
  int o_tile, o_red, o_region;
  int o_zero;                      // 8 rows of zeros at the widest activation stride: operands of padding tiles
  int tiles_j1, tiles_1, tiles_jl, tiles_l;  // dW tiles of the first layer (two-layer policies) and of the last layer
};

// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): g = lane >> 2, t = lane & 3;
// A (8x4, row) a = A[g][t], B (4x8, col) b = B[t][g], C (8x8) c0, c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// C[g][n] = sum_k X[g][k] B[k][n] for the 8 rollout rows g, n < n_cols: output tiles of 8 columns dealt over the
// warps, two tiles per trip and two accumulator chains per tile so that four DMMAs are in flight (a dependent
// DMMA issues every ~26 cycles, scripts/microbench/fp64_peak.cu).  B[k][n] = W[k * wk + n * wn]: the forward
// products read the weights as stored (wk = row stride, wn = 1), the input adjoints read them transposed.
// epi(g, n, value) receives every element of the tile.
template <class Epi>
__device__ __forceinline__ void denseTiles(const double *X, int sX, const double *W, int wk, int wn, int k_pad, int n_cols,
                                           int warp, int n_warps, int lane, Epi epi) {
  const int g = lane >> 2, t = lane & 3;
  const double *xa = X + g * sX + t;
  for (int n0 = warp * 8; n0 < n_cols; n0 += 2 * n_warps * 8) {
    const int n1 = n0 + n_warps * 8;
    const bool two = n1 < n_cols;
    const double *wa = W + t * wk + (n0 + g) * wn, *wb = W + t * wk + (n1 + g) * wn;
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    int k = 0;
    if (two) {
      for (; k + 8 <= k_pad; k += 8) {
        const double x0 = xa[k], x1 = xa[k + 4];
        dmma884(a0, a1, x0, wa[k * wk]);
        dmma884(c0, c1, x0, wb[k * wk]);
        dmma884(b0, b1, x1, wa[(k + 4) * wk]);
        dmma884(d0, d1, x1, wb[(k + 4) * wk]);
      }
      if (k < k_pad) {
        const double x0 = xa[k];
        dmma884(a0, a1, x0, wa[k * wk]);
        dmma884(c0, c1, x0, wb[k * wk]);
      }
      epi(g, n1 + 2 * t, c0 + d0);
      epi(g, n1 + 2 * t + 1, c1 + d1);
    } else {
      for (; k + 8 <= k_pad; k += 8) {
        dmma884(a0, a1, xa[k], wa[k * wk]);
        dmma884(b0, b1, xa[k + 4], wa[(k + 4) * wk]);
      }
      if (k < k_pad) dmma884(a0, a1, xa[k], wa[k * wk]);
    }
    epi(g, n0 + 2 * t, a0 + b0);
    epi(g, n0 + 2 * t + 1, a1 + b1);
  }
}

// development aid: -DTRX_ROLLOUT_PROFILE accumulates clock64() per phase (thread 0 of every CTA, CTA 0 reported)
#ifdef TRX_ROLLOUT_PROFILE
#define PH(idx, ...)                                  \
  {                                                   \
    long long _t0 = clock64();                        \
    __VA_ARGS__;                                      \
    if (tid == 0) sprof[idx] += clock64() - _t0;      \
  }
__device__ long long g_rollout_prof[64];
#else
#define PH(idx, ...) \
  { __VA_ARGS__; }
#endif

// lane exchange of the warp-scan phases (rollout_frame.h fkRound*): shuffles, component by component
struct DeviceWarp {
  template <class T> static __device__ __forceinline__ void up(const T (&x)[1], int d, T (&o)[1]) {
    const double *src = reinterpret_cast<const double *>(&x[0]);
    double *dst = reinterpret_cast<double *>(&o[0]);
#pragma unroll
    for (int e = 0; e < (int)(sizeof(T) / sizeof(double)); e++) dst[e] = __shfl_up_sync(0xffffffffu, src[e], d);
  }
  template <class T> static __device__ __forceinline__ void down(const T (&x)[1], int d, T (&o)[1]) {
    const double *src = reinterpret_cast<const double *>(&x[0]);
    double *dst = reinterpret_cast<double *>(&o[0]);
#pragma unroll
    for (int e = 0; e < (int)(sizeof(T) / sizeof(double)); e++) dst[e] = __shfl_down_sync(0xffffffffu, src[e], d);
  }
};

// block barriers: all warps of the CTA, or the NW main warps only (named barrier 1)
__device__ __forceinline__ void barAll() { __syncthreads(); }
template <int NTHREADS> __device__ __forceinline__ void barMain() {
  asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

// 8-byte asynchronous copy global -> shared (LDGSTS)
__device__ __forceinline__ void cpAsync8(double *dst_shared, const double *src_global) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global)
               : "memory");
}

// 16-byte asynchronous copy global -> shared (both ends 16-byte aligned)
__device__ __forceinline__ void cpAsync16(double *dst_shared, const double *src_global) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global)
               : "memory");
}

constexpr int kHelperWarps = 8;            // helper warps of a CTA:
constexpr int kDwWarps = kHelperWarps - 1;  // weight-gradient warps, and one body warp

// One CTA = one group of NW rollouts at a time, NW main warps + 8 helper warps:
//   main warps     the frame phases (rollout_frame.h), the policy layers and their input adjoints (DMMA tiles)
//   helper warps   what is off the frame's critical path: the body step and its adjoint (helper warp 0, a lane per
//                  rollout) while the main warps run the forward kinematics, and the weight gradients dW = X^T GY
//                  (register accumulators: U1 tiles of the first layer of a two-layer policy and UL tiles of the
//                  last layer per helper warp, over all frames and groups of the CTA) plus the bias gradients
//                  while the main warps go on with the adjoint of the frame.  In the forward sweep, where there is
//                  no dW, the same warps write the frame's checkpoint beyond the carried state (full checkpoints).
// The two sides meet at the full barriers of a step (forward: A, B, S; adjoint: A..G); everything else synchronises
// the main warps only.
// kPlain: the evaluation stores no residual vector and takes no adjoint seeds (loss and g = J^T r only -- the solver
// steps): compiled without the residual-row addressing of the general form
// kFull: full checkpoints (the default mode) known at compile time -- the recompute paths of the other modes drop out
template <int NW, int U1, int UL, bool kPlain, bool kFull>
__global__ void __launch_bounds__((NW + kHelperWarps) * 32, 1)
    trx_rollout_kernel(const __grid_constant__ KHeader hdr, const __grid_constant__ KArgs a) {
  double *const a_r_out = kPlain ? nullptr : a.r_out;
  const double *const a_seed = kPlain ? nullptr : a.seed;
  extern __shared__ double sm[];
  const int32_t *H = hdr.v;
  constexpr int NH = kHelperWarps, NHD = kDwWarps, NTM = NW * 32, NTH = NH * 32, NTD = NHD * 32, MAXT = U1 + UL;
  const int tid = threadIdx.x, nt = NTM + NTH, warp = tid >> 5, lane = tid & 31;
#ifdef TRX_ROLLOUT_PROFILE
  __shared__ long long sprof[64];
  if (tid < 64) sprof[tid] = 0;
#endif
  const int n_in = H[H_NIN], n_hid = H[H_NHID], n_out = H[H_NOUT], n_layers = H[H_NLAYERS];
  const int n1 = n_layers == 2 ? n_hid : n_out;  // output width of the first layer

  // ---- stage the plan and the weights
  double *sD = sm + a.o_D;
  int32_t *sI = reinterpret_cast<int32_t *>(sm + a.o_I);
  for (int e = tid; e < a.nd; e += nt) sD[e] = a.D[e];
  for (int e = tid; e < a.ni; e += nt) sI[e] = a.I[e];
  for (int e = tid; e < a.o_region - a.o_W1; e += nt) sm[a.o_W1 + e] = 0.0;  // weights (padded), activations, tables
  __syncthreads();
  if (n_layers) {
    const double *W1 = a.x + H[H_W1], *B1 = a.x + H[H_B1];
    for (int e = tid; e < n_in * n1; e += nt) {
      int i = e / n1, j = e - i * n1;
      sm[a.o_W1 + i * a.s_W1 + j] = W1[e];
    }
    for (int e = tid; e < n1; e += nt) sm[a.o_B1 + e] = B1[e];
    if (n_layers == 2) {
      const double *W2 = a.x + H[H_W2], *B2 = a.x + H[H_B2];
      for (int e = tid; e < n_hid * n_out; e += nt) {
        int i = e / n_out, j = e - i * n_out;
        sm[a.o_W2 + i * a.s_W2 + j] = W2[e];
      }
      for (int e = tid; e < n_out; e += nt) sm[a.o_B2 + e] = B2[e];
    }
  }
  // dW tile table, [dW warp][U1 + UL] entries {A operand column block, B operand column block} (offsets in
  // doubles): tile idx = dW warp + u * NHD of its layer; slots past the layer's tile count point both operands
  // at zeros, so the accumulation is straight-line code without a test per tile
  int2 *tile_tab = reinterpret_cast<int2 *>(sm + a.o_tile);
  for (int e = tid; e < NHD * MAXT; e += nt) {
    const int w = e / MAXT, u = e - w * MAXT;
    int2 ent = make_int2(a.o_zero, a.o_zero);
    if (u < U1) {
      const int idx = w + u * NHD;
      if (n_layers == 2 && idx < a.tiles_1) ent = make_int2(a.o_X + (idx / a.tiles_j1) * 8, a.o_GH + (idx % a.tiles_j1) * 8);
    } else {
      const int idx = w + (u - U1) * NHD;
      if (idx < a.tiles_l)
        ent = make_int2((n_layers == 2 ? a.o_A : a.o_X) + (idx / a.tiles_jl) * 8, a.o_GO + (idx % a.tiles_jl) * 8);
    }
    tile_tab[e] = ent;
  }
  __syncthreads();

  const double *sW1 = sm + a.o_W1, *sB1 = sm + a.o_B1, *sW2 = sm + a.o_W2, *sB2 = sm + a.o_B2;
  double *XS = sm + a.o_X, *HS = sm + a.o_H, *AS = sm + a.o_A, *OS = sm + a.o_O;
  double *GXS = sm + a.o_GX, *GHS = sm + a.o_GH, *GOS = sm + a.o_GO;
  // a frame's checkpoint: the carried state (H_CKPT doubles), then what the policy layers produced in the frame --
  // hidden pre-activations, activations, outputs, tanh of the joint commands -- so that the adjoint sweep recomputes
  // kinematics and contacts only
  const int T = a.frames, CK = H[H_CKPT];
  const int CKX = a.ckpt_doubles;  // stride
  const int mode = kFull ? 2 : a.ckpt_mode;
  // offsets of the optional parts inside a checkpoint
  // (every segment starts on a 16-byte boundary and is an even number of doubles: bulk copies move them whole)
  const int ev = 1;  // round up to even
  const int CKP = (CK + ev) & ~ev, seg_h = (n_hid + ev) & ~ev, seg_o = (n_out + ev) & ~ev, seg_v = (H[H_NQ] + ev) & ~ev;
  const int seg_pose = (H[H_NJ] * 7 + ev) & ~ev, seg_x = (n_in + ev) & ~ev, seg_ee = H[H_NEE] * 12;
  const int CK_POLICY = CKP, CK_FULL = CKP + 2 * seg_h + seg_o + seg_v;
  const int CK_X = CK_FULL + seg_pose, CK_EE = CK_X + seg_x, CK_CG = CK_EE + seg_ee;
  const int n_steps = a.want_grad ? 2 * T : T;
  double *part = a.partial + (size_t)blockIdx.x * (1 + a.nx);

  // Full checkpoints (mode 2): the adjoint sweep recomputes nothing.  A frame's checkpoint -- carried state, policy
  // values, link poses, policy input, contact points and velocities, body centre -- comes back with 16-byte
  // asynchronous copies into the places the forward sweep had it in, each part as soon as the adjoint of the frame
  // before has read that place for the last time, so that the copies fly while that adjoint is still running:
  //   part 1 (after barrier C): policy outputs, commands, contact points / velocities, body centre
  //   part 2 (after D): hidden pre-activations and activations     part 4 (after E): policy input
  //   part 8 (top of the frame's own step, main warps only): carried state and link poses, first read by the
  //          second contact phase
  // thread = (rollout t & 7, 16-byte unit t >> 3); one cp.async group per call.
  auto fetchCheckpoint = [&](int grp, int f, int parts, int n_threads) {
    const int rr = tid & (NW - 1), u0 = tid / NW, us = n_threads / NW;
    const int live = (int)min((long)NW, a.count - (long)grp * NW);
    if (rr < live && f >= 0 && tid < n_threads) {
      const double *ck = a.ckpt + (((size_t)grp * NW + rr) * T + f) * CKX;
      double *reg = sm + a.o_region + rr * H[H_REGION];
      auto seg = [&](double *dst, const double *src, int n) {
        for (int u = u0; 2 * u < n; u += us) cpAsync16(dst + 2 * u, src + 2 * u);
      };
      if (parts & 1) {
        seg(OS + rr * a.s_O, ck + CK_POLICY + 2 * seg_h, seg_o);
        seg(reg + H[R_VEL], ck + CK_POLICY + 2 * seg_h + seg_o, seg_v);
        for (int e = 0; e < H[H_NEE]; e++) seg(reg + H[R_EE] + e * EE_STRIDE, ck + CK_EE + e * 12, 12);
        seg(reg + H[R_CG], ck + CK_CG, 4);
      }
      if (parts & 2) {
        seg(HS + rr * a.s_H, ck + CK_POLICY, seg_h);
        seg(AS + rr * a.s_H, ck + CK_POLICY + seg_h, seg_h);
      }
      if (parts & 4) seg(XS + rr * a.s_X, ck + CK_X, seg_x);
      if (parts & 8) {
        seg(reg + H[R_Q], ck, CKP);
        seg(reg + H[R_POSE], ck + CK_FULL, seg_pose);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto fetchWait = [&](int pending) {
    if (pending == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.wait_group 1;" ::: "memory");
  };

  if (warp == NW + NHD) {
    // ================================================================================ body warp
    // lane = rollout of the group: the body step while the main warps run the forward kinematics, its adjoint
    // while they run the adjoint of the forward kinematics
    Ctx cb;
    cb.H = H, cb.I = sI, cb.D = sD;
    cb.s = sm + a.o_region + (lane & (NW - 1)) * H[H_REGION];
    cb.frames = T;
    for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
      const bool body_lane = lane < (int)min((long)NW, a.count - (long)grp * NW);
      barAll();  // the group starts
      for (int step = 0; step <= n_steps; step++) {
        if ((step == T && !a.want_grad) || step == n_steps) break;
        if (step >= T && mode == 2) {
          if (step == T) {
            barAll();  // F: the main warps are done with the last forward frame
            fetchCheckpoint(grp, T - 1, 7, nt);
          }
          fetchWait(0);
        }
        barAll();  // A: the frame's carried state is in place
        if (step >= T && mode == 2) {
          // nothing to recompute: pull the next frame's checkpoints of the group towards L2 instead
          const int f = 2 * T - 1 - step;
          if (f > 0)
            for (int rr = 0; rr < NW; rr++) {
              const long local = (long)grp * NW + rr;
              if (local >= a.count) break;
              const double *ck = a.ckpt + ((size_t)local * T + (f - 1)) * CKX;
              for (int line = lane; line * 16 < CKX; line += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(ck + line * 16));
            }
        } else if (body_lane) {
          phaseBody(cb);
        }
        barAll();  // B
        if (step < T) {
          if (mode == 2) barAll();  // S
          continue;
        }
        const int f_next = 2 * T - 2 - step;  // the frame of the next step
        barAll();  // C
        if (mode == 2) fetchCheckpoint(grp, f_next, 1, nt);
        barAll();  // D
        if (mode == 2) fetchCheckpoint(grp, f_next, 2, nt);
        barAll();  // E: the object's pose adjoint is complete
        if (mode == 2) fetchCheckpoint(grp, f_next, 4, nt);
        if (body_lane) bphaseBody(cb);
        barAll();  // G: the frame is done
      }
    }
    return;
  }
  if (warp >= NW) {
    // ================================================================================ dW warps
    const int hw = warp - NW, htid = tid - NTM;
    const int gi = lane >> 2, t4 = lane & 3;
    double acc[MAXT][2];
#pragma unroll
    for (int u = 0; u < MAXT; u++) acc[u][0] = acc[u][1] = 0.0;
    double gb0 = 0.0, gb1 = 0.0;  // bias gradients: element htid and htid + NTD of {b1, b2}
    // lane parts of the dW operand addresses: A[g][t] = X[rollout t][column g], B[t][g] = GY[rollout t][column g]
    const int s_xl = n_layers == 2 ? a.s_H : a.s_X;  // last layer: X = activations of the hidden layer
    const int2 *my_tiles = tile_tab + hw * MAXT;
    for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
      barAll();  // the group starts
      for (int step = 0; step <= n_steps; step++) {
        if ((step == T && !a.want_grad) || step == n_steps) break;
        if (step >= T && mode == 2) {
          if (step == T) {
            barAll();  // F
            fetchCheckpoint(grp, T - 1, 7, nt);
          }
          fetchWait(0);
        }
        barAll();  // A
        barAll();  // B
        if (step < T) {
          if (mode == 2) {
            barAll();  // S: the frame's values are final; the main warps go on with the last contact phase
            // the frame's checkpoint beyond the carried state: thread = (rollout, slot) over the dW warps
            const int rr = htid & (NW - 1), sl = htid / NW;
            constexpr int ns = NTD / NW;
            const long local = (long)grp * NW + rr;
            if (local < a.count) {
              double *ck = a.ckpt + ((size_t)local * T + step) * CKX;
              const double *reg = sm + a.o_region + rr * H[H_REGION];
              const double *hr = HS + rr * a.s_H, *ar = AS + rr * a.s_H, *outr = OS + rr * a.s_O, *xr = XS + rr * a.s_X;
              const int n_pose = H[H_NJ] * 7, nqv = H[H_NQ];
              double *cp = ck + CK_POLICY;
              for (int e = sl; e < seg_h; e += ns) cp[e] = e < n_hid ? hr[e] : 0.0, cp[seg_h + e] = e < n_hid ? ar[e] : 0.0;
              for (int e = sl; e < seg_o; e += ns) cp[2 * seg_h + e] = e < n_out ? outr[e] : 0.0;
              for (int e = sl; e < seg_v; e += ns) cp[2 * seg_h + seg_o + e] = e < nqv ? reg[H[R_VEL] + e] : 0.0;
              for (int e = sl; e < seg_pose; e += ns) ck[CK_FULL + e] = e < n_pose ? reg[H[R_POSE] + e] : 0.0;
              for (int e = sl; e < seg_x; e += ns) ck[CK_X + e] = e < n_in ? xr[e] : 0.0;
              for (int e = sl; e < seg_ee; e += ns) ck[CK_EE + e] = reg[H[R_EE] + (e / 12) * EE_STRIDE + e % 12];
              for (int e = sl; e < 4; e += ns) ck[CK_CG + e] = e < 3 ? reg[H[R_CG] + e] : 0.0;
            }
          }
          continue;
        }
        const int f_next = 2 * T - 2 - step;  // the frame of the next step
        barAll();  // C: output adjoint of the last layer complete
        if (mode == 2) fetchCheckpoint(grp, f_next, 1, nt);
#ifdef TRX_ROLLOUT_PROFILE
        long long _h0 = clock64();
#endif
        {
          // dW += X^T GY, K = the 8 rollouts = two DMMAs per 8x8 tile; the first DMMA of every tile, then the
          // second, so that consecutive DMMAs never wait on each other's accumulator
          const double *xl = sm + gi + t4 * s_xl, *yl = sm + gi + t4 * a.s_O;
#pragma unroll
          for (int half = 0; half < 2; half++) {
#pragma unroll
            for (int u = U1; u < MAXT; u++) {
              const int2 tt = my_tiles[u];
              dmma884(acc[u][0], acc[u][1], xl[tt.x + 4 * half * s_xl], yl[tt.y + 4 * half * a.s_O]);
            }
          }
          const int base = n_layers == 2 ? n_hid : 0;
          for (int e = htid, sl2 = 0; e < base + n_out; e += NTD, sl2++) {
            if (e < base) continue;
            const double *gy = GOS + (e - base);
            double sgy = 0;
            for (int g = 0; g < 8; g++) sgy += gy[g * a.s_O];
            if (sl2 == 0) gb0 += sgy;
            else gb1 += sgy;
          }
        }
#ifdef TRX_ROLLOUT_PROFILE
        if (htid == 0) sprof[50] += clock64() - _h0;
#endif
        barAll();  // D: output adjoint of the first layer complete
#ifdef TRX_ROLLOUT_PROFILE
        _h0 = clock64();
#endif
        if (n_layers == 2) {
          if (U1 > 0) {
            const double *x1 = sm + gi + t4 * a.s_X, *y1 = sm + gi + t4 * a.s_H;
#pragma unroll
            for (int half = 0; half < 2; half++) {
#pragma unroll
              for (int u = 0; u < U1; u++) {
                const int2 tt = my_tiles[u];
                dmma884(acc[u][0], acc[u][1], x1[tt.x + 4 * half * a.s_X], y1[tt.y + 4 * half * a.s_H]);
              }
            }
          }
          for (int e = htid, sl2 = 0; e < n_hid; e += NTD, sl2++) {
            const double *gy = GHS + e;
            double sgy = 0;
            for (int g = 0; g < 8; g++) sgy += gy[g * a.s_H];
            if (sl2 == 0) gb0 += sgy;
            else gb1 += sgy;
          }
        }
#ifdef TRX_ROLLOUT_PROFILE
        if (htid == 0) sprof[51] += clock64() - _h0;
#endif
        if (mode == 2) fetchCheckpoint(grp, f_next, 2, nt);
        barAll();  // E
        if (mode == 2) fetchCheckpoint(grp, f_next, 4, nt);
        barAll();  // G: the frame is done
      }
    }
    // ---- this CTA's partial gradient
    if (a.want_grad && n_layers) {
#pragma unroll
      for (int u = 0; u < MAXT; u++) {
        int ti, tj, base, width, rows;
        if (u < U1) {
          const int idx = hw + u * NHD;
          if (n_layers != 2 || idx >= a.tiles_1) continue;
          ti = idx / a.tiles_j1, tj = idx - ti * a.tiles_j1;
          base = H[H_W1], width = n1, rows = n_in;
        } else {
          const int idx = hw + (u - U1) * NHD;
          if (idx >= a.tiles_l) continue;
          ti = idx / a.tiles_jl, tj = idx - ti * a.tiles_jl;
          base = n_layers == 2 ? H[H_W2] : H[H_W1], width = n_out, rows = n_layers == 2 ? n_hid : n_in;
        }
        const int i = ti * 8 + gi, j = tj * 8 + 2 * t4;
        if (i < rows) {
          double *p = part + 1 + base + i * width + j;
          if (j < width) p[0] = a.accumulate ? p[0] + acc[u][0] : acc[u][0];
          if (j + 1 < width) p[1] = a.accumulate ? p[1] + acc[u][1] : acc[u][1];
        }
      }
      const int nb = (n_layers == 2 ? n_hid : 0) + n_out;
      for (int e = htid, sl2 = 0; e < nb; e += NTD, sl2++) {
        const double v = sl2 == 0 ? gb0 : gb1;
        // {b1, b2} order; with one layer the only bias is b1
        const int idx = n_layers == 2 ? (e < n_hid ? H[H_B1] + e : H[H_B2] + (e - n_hid)) : H[H_B1] + e;
        part[1 + idx] = a.accumulate ? part[1 + idx] + v : v;
      }
    }
    return;
  }

  // ================================================================================== main warps
  const int k1 = (n_in + 3) & ~3, k2 = (n_hid + 3) & ~3;
  // Work mapping outside the policy layers: thread = (rollout r of the CTA's group, slot); a phase with n items runs
  // item = slot, slot + 32, ... of rollout r, so the lanes of a warp hold the SAME items of all NW rollouts.  Phases
  // with few items (a dozen contact sides, six end effectors) then fill their warps across rollouts instead of
  // issuing FP64 instructions for a few lanes per rollout (an FP64 warp instruction occupies the pipe for two cycles
  // whatever its lane mask), and a phase's control flow is uniform across the lanes of an item.
  const int r = tid & (NW - 1), slot = tid / NW;
  Ctx c;
  c.H = H;
  c.I = sI;
  c.D = sD;
  c.s = sm + a.o_region + r * H[H_REGION];
  c.x = XS + r * a.s_X, c.h = HS + r * a.s_H, c.a = AS + r * a.s_H, c.out = OS + r * a.s_O;
  c.gx = GXS + r * a.s_X, c.gh = GHS + r * a.s_H, c.gout = GOS + r * a.s_O;
  c.r_stride = a.R;
  c.frames = a.frames;
  // The forward kinematics and its adjoint run with the other mapping: warp = rollout, lane = joint of a scan
  // round (rollout_frame.h fkRound*); no block barrier inside them.
  Ctx cw = c;
  cw.s = sm + a.o_region + warp * H[H_REGION];
  const int n_round = H[H_NROUND];

  double loss = 0.0;
  const double w_reg = sD[H[H_D_CONST] + DC_W_REG], w_creg = sD[H[H_D_CONST] + DC_W_CREG];
  const int nq = H[H_NQ], rows_frame = H[H_ROWS_FRAME];
  constexpr int kPre = 6;  // checkpoint doubles per thread (H_CKPT <= 192, checked by the host)
  double pre[kPre];

  for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
    const long grp_first = (long)grp * NW;  // first rollout of the group inside this launch
    const int live = (int)min((long)NW, a.count - grp_first);  // rows of the group that are rollouts
    const long local = grp_first + r;
    const bool active = r < live;
    const long rollout = a.first + local;
    double *ckpt = a.ckpt + (size_t)local * T * CKX;
    // element (row, g) of the group's residual rows / seeds at [row * R + g]
    double *grp_r = a_r_out ? a_r_out + H[H_NSCALAR] + a.first + grp_first : nullptr;
    const double *grp_seed = a_seed ? a_seed + H[H_NSCALAR] + a.first + grp_first : nullptr;
    c.r_out = (a_r_out && active) ? a_r_out + H[H_NSCALAR] + rollout : nullptr;
    c.seed = nullptr;
    barAll();  // the group starts: everybody is done with the previous one
    if (!active) {
      // an idle row of the last group: its rows of the activation matrices must stay zero
      for (int e = slot; e < a.s_X; e += 32) c.x[e] = 0.0, c.gx[e] = 0.0;
      for (int e = slot; e < a.s_H; e += 32) c.h[e] = 0.0, c.a[e] = 0.0, c.gh[e] = 0.0;
      for (int e = slot; e < a.s_O; e += 32) c.out[e] = 0.0, c.gout[e] = 0.0;
    } else {
      const double pr[3] = {a.params[rollout], a.params[a.R + rollout], a.params[2 * a.R + rollout]};
      phaseInit(c, slot, pr);
    }
    barMain<NTM>();
    if (warp < live)
      for (int rd = 0; rd < n_round; rd++) {
        fkRoundForward<1, DeviceWarp>(cw, rd, lane);
        __syncwarp();
      }
    barMain<NTM>();
    // One loop over both sweeps so that the frame's forward code exists once: steps 0 .. T-1 are the forward
    // sweep (frame = step), steps T .. 2T-1 the adjoint sweep (frame = 2T-1-step: recompute it from its
    // checkpoint, then its adjoint).
    for (int step = 0; step <= n_steps; step++) {
      const bool rec = step >= T;
      const int f = rec ? 2 * T - 1 - step : step;
      if (step == T) {
        // between the sweeps: terminal goals, then the adjoint's carried state
        if (active) phaseTerminal(c, slot, loss);
        if (!a.want_grad) break;
        c.r_out = nullptr;
        c.seed = (a_seed && active) ? a_seed + H[H_NSCALAR] + rollout : nullptr;
        if (active) {
          bphaseInit(c, slot);
          if (mode != 2) {
            const double *ck = ckpt + (size_t)(T - 1) * CKX;
#pragma unroll
            for (int i = 0; i < kPre; i++) pre[i] = (slot + 32 * i < CK) ? ck[slot + 32 * i] : 0.0;
          }
        }
        if (mode == 2) barAll();  // F: the body warp may overwrite the forward values with the checkpoint's
        else barMain<NTM>();
      }
      if (step == n_steps) break;

      // ---- the frame, forward
      if (!rec) {
        PH(12, if (active) phaseBeginFrame(c, slot, f, ckpt + (size_t)f * CKX));
      } else {
        PH(13, if (active && mode != 2) {
          phaseLoadCheckpoint(c, slot, pre);
          const double *ck = ckpt + (size_t)f * CKX;
          if (mode >= 1) {
            // the frame's policy values travel from the checkpoint straight into the activation matrices while
            // the kinematics are recomputed (nobody reads them before the contact phases)
            const double *cp = ck + CK_POLICY;
            for (int e = slot; e < n_hid; e += 32) {
              cpAsync8(c.h + e, cp + e);
              cpAsync8(c.a + e, cp + seg_h + e);
            }
            for (int e = slot; e < n_out; e += 32) cpAsync8(c.out + e, cp + 2 * seg_h + e);
            for (int e = slot; e < nq; e += 32) cpAsync8(c.s + H[R_VEL] + e, cp + 2 * seg_h + seg_o + e);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        });
      }
      if (rec && mode == 2) {
        // carried state and link poses of this frame: nobody needs them before the second contact phase
        PH(13, if (step == T) fetchCheckpoint(grp, f, 7, nt); fetchCheckpoint(grp, f, 8, NTM); fetchWait(1));
      }
      barAll();  // A: the body step may start
      if (rec && mode != 2 && active && f > 0) {  // the next checkpoint lands while this frame is being worked on
        const double *ck = ckpt + (size_t)(f - 1) * CKX;
#pragma unroll
        for (int i = 0; i < kPre; i++) pre[i] = (slot + 32 * i < CK) ? ck[slot + 32 * i] : 0.0;
      }
      if (rec && mode == 2) {
        // nothing is recomputed: the joints advance (q += command * dt), the body state the frame started from is
        // the carried one, the rest arrives from the checkpoint
        PH(15, barAll());  // B
      } else {
        PH(0, if (active) phaseSim(c, slot); barMain<NTM>());
        PH(1, if (warp < live) for (int rd = 0; rd < n_round; rd++) {
          fkRoundForward<1, DeviceWarp>(cw, rd, lane);
          __syncwarp();
        });
        PH(15, barAll());  // B: link poses and the body step are complete
        PH(2, if (active) { if (rec) phaseObserve<false>(c, slot, f, loss); else phaseObserve<true>(c, slot, f, loss); });
        if (n_layers) {
          // The policy layers' epilogues carry what follows them: activity / command regularisation rows and
          // tanh (dexenv_grasp.h:137-146,186-199), so the hidden and control steps cost no phase of their own.
          const long row_act = (long)f * rows_frame + H[H_ROW_ACT], row_cmd = (long)f * rows_frame + H[H_ROW_CMD];
          const bool layers_now = !rec || mode == 0;  // otherwise their values come from the checkpoint
          const bool count = !rec;                    // residual rows count in the forward sweep only
          if (rec) asm volatile("cp.async.wait_group 0;" ::: "memory");
          PH(3, barMain<NTM>());
          if (layers_now && n_layers == 2) {
            PH(4, denseTiles(XS, a.s_X, sW1, a.s_W1, 1, k1, n_hid, warp, NW, lane, [=, &loss](int g, int j, double v) {
              v += sB1[j];
              HS[g * a.s_H + j] = v;
              AS[g * a.s_H + j] = tanh(v);
              if (count && j < n_hid && g < live) {
                const double rr = v * w_reg;
                loss += rr * rr;
                if (grp_r) grp_r[(row_act + j) * a.R + g] = rr;
              }
            }));
            PH(3, barMain<NTM>());
          }
          if (layers_now) {
            const double *Xl = n_layers == 2 ? AS : XS, *Wl = n_layers == 2 ? sW2 : sW1, *Bl = n_layers == 2 ? sB2 : sB1;
            const int sXl = n_layers == 2 ? a.s_H : a.s_X, sWl = n_layers == 2 ? a.s_W2 : a.s_W1;
            double *vel = sm + a.o_region + H[R_VEL];
            const int region = H[H_REGION];
            PH(6, denseTiles(Xl, sXl, Wl, sWl, 1, n_layers == 2 ? k2 : k1, n_out, warp, NW, lane,
                             [=, &loss](int g, int j, double v) {
                               v += Bl[j];
                               OS[g * a.s_O + j] = v;
                               if (j < nq && g < NW) vel[g * region + j] = tanh(v);
                               if (count && j < n_out && g < live) {
                                 const double rr = v * w_creg;
                                 loss += rr * rr;
                                 if (grp_r) grp_r[(row_cmd + j) * a.R + g] = rr;
                               }
                             }));
            PH(3, barMain<NTM>());
            if (!rec && mode == 1 && active) {
              double *cp = ckpt + (size_t)f * CKX + CK_POLICY;
              for (int e = slot; e < seg_h; e += 32) cp[e] = e < n_hid ? c.h[e] : 0.0, cp[seg_h + e] = e < n_hid ? c.a[e] : 0.0;
              for (int e = slot; e < seg_o; e += 32) cp[2 * seg_h + e] = e < n_out ? c.out[e] : 0.0;
              for (int e = slot; e < seg_v; e += 32) cp[2 * seg_h + seg_o + e] = e < nq ? c.s[H[R_VEL] + e] : 0.0;
            }
          }
          PH(8, if (active) { if (rec) phaseContact1<false>(c, slot, f, loss); else phaseContact1<true>(c, slot, f, loss); } barMain<NTM>());
          // S (full checkpoints, forward sweep): everything a checkpoint holds beyond the carried state is final -- the dW
          // warps, idle in this sweep, copy it out while the main warps finish the frame
          PH(9, if (active) phaseContact2(c, slot, f); if (!rec && mode == 2) barAll(); else barMain<NTM>());
          if (!rec) {
            PH(10, if (active) phaseContact3(c, slot, f, loss); barMain<NTM>());
          }
        }
      }
      if (!rec) continue;

      // ---- the frame, adjoint
      PH(32, if (active) bphaseBegin(c, slot, f); barMain<NTM>(); if (f == T - 1 || f == 0) {
        if (active) bphaseBegin2(c, slot, f);
        barMain<NTM>();
      });
      if (n_layers) {
        PH(33, if (active) bphaseContact3(c, slot, f); if (mode == 2) fetchWait(0); barMain<NTM>());
        if (mode == 2) {
          // the carried state has landed: the joints advance, the body state the frame started from is the carried one
          PH(0, if (active) phaseAdvance(c, slot));
        }
        PH(34, if (active) bphaseContact2(c, slot, f); barMain<NTM>());
        PH(35, if (active) bphaseContact1(c, slot, f); barMain<NTM>());
        PH(36, if (active) bphaseControl(c, slot, f));
      }
      const int f_next = f - 1;
      PH(46, barAll());  // C: the helper warps take dW and db of the last layer from here
      if (mode == 2) fetchCheckpoint(grp, f_next, 1, nt);
      if (n_layers == 2) {
        // input adjoint of the last layer, with the tanh / activity-row adjoint of the hidden layer in its epilogue
        const long row_act = (long)f * rows_frame + H[H_ROW_ACT];
        PH(38, denseTiles(GOS, a.s_O, sW2, 1, a.s_W2, (n_out + 3) & ~3, n_hid, warp, NW, lane,
                          [=](int g, int i, double v) {
                            const bool on = g < live && i < n_hid;
                            const double av = AS[g * a.s_H + i];
                            const double sd = grp_seed ? (on ? grp_seed[(row_act + i) * a.R + g] : 0.0)
                                                       : HS[g * a.s_H + i] * w_reg;
                            GHS[g * a.s_H + i] = on ? v * (1.0 - av * av) + sd * w_reg : 0.0;
                          }));
      } else if (n_layers == 1) {
        PH(38, denseTiles(GOS, a.s_O, sW1, 1, a.s_W1, (n_out + 3) & ~3, n_in, warp, NW, lane,
                          [=](int g, int i, double v) { GXS[g * a.s_X + i] = v; }));
      }
      PH(47, barAll());  // D: ... and of the first layer from here
      if (mode == 2) fetchCheckpoint(grp, f_next, 2, nt);
      if (n_layers == 2) {
        PH(41, denseTiles(GHS, a.s_H, sW1, 1, a.s_W1, k2, n_in, warp, NW, lane,
                          [=](int g, int i, double v) { GXS[g * a.s_X + i] = v; }));
        PH(37, barMain<NTM>());
      }
      PH(43, if (active) bphaseObserve(c, slot, f); barMain<NTM>(); if (warp < live) bphaseObjectGather(cw, lane));
      PH(48, barAll());  // E: the body adjoint may start
      if (mode == 2) fetchCheckpoint(grp, f_next, 4, nt);
      PH(44, if (warp < live) {
        for (int rd = n_round - 1; rd >= 0; rd--) {
          fkRoundReverse<1, DeviceWarp>(cw, rd, lane);
          __syncwarp();
        }
        fkClearStatics(cw, lane);
        bphaseSim(cw, lane);
      });
      PH(49, barAll());  // G
    }
  }

#ifdef TRX_ROLLOUT_PROFILE
  barMain<NTM>();
  if (blockIdx.x == 0 && tid < 64) g_rollout_prof[tid] = sprof[tid];
#endif
  // ---- this CTA's partial loss
  {
    for (int off = 16; off > 0; off >>= 1) loss += __shfl_down_sync(0xffffffffu, loss, off);
    double *red = sm + a.o_red;
    barMain<NTM>();
    if (lane == 0) red[warp] = loss;
    barMain<NTM>();
    if (tid == 0) {
      double sum = 0;
      for (int w = 0; w < NW; w++) sum += red[w];
      part[0] = a.accumulate ? part[0] + sum : sum;
    }
  }
}

// deterministic cross-CTA sum: out[0] = loss, out[1 + i] = gradient; the lane-independent weight-decay rows
// (neural.h:159-171: r = reg * w) are added here, once, when this rank owns them
__global__ void trx_rollout_reduce_kernel(const double *__restrict__ partial, int n_parts, int nx,
                                          const double *__restrict__ x, double reg, int scalar_rows, int want_grad,
                                          const double *__restrict__ seed, double *__restrict__ r_out,
                                          double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (want_grad && i < nx) {
    double s = 0;
    for (int p = 0; p < n_parts; p++) s += partial[(size_t)p * (1 + nx) + 1 + i];
    if (scalar_rows) s += (seed ? seed[i] : x[i] * reg) * reg;
    out[1 + i] = s;
  }
  if (r_out && scalar_rows && i < nx) r_out[i] = x[i] * reg;
  if (blockIdx.x == 0) {
    __shared__ double red[256];
    double s = 0;
    if (scalar_rows)
      for (int k = threadIdx.x; k < nx; k += blockDim.x) {
        double v = x[k] * reg;
        s += v * v;
      }
    for (int p = threadIdx.x; p < n_parts; p += blockDim.x) s += partial[(size_t)p * (1 + nx)];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
      if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
  }
}

}  // namespace

// ------------------------------------------------------------------ host side

struct trx_rollout {
  Plan plan;
  KHeader hdr;
  KArgs args;
  int device = 0;
  int nw = 8, u1 = 0, ul = 0, grid = 0;
  size_t smem_bytes = 0;
  uint64_t lanes = 0, chunk = 0, n_chunks = 0;
  bool scalar_rows = true;
  int32_t *d_I = nullptr;
  double *d_D = nullptr, *d_params = nullptr, *d_ckpt = nullptr, *d_partial = nullptr;
  uint64_t launches = 0;
  double reg = 0;
};

namespace {

typedef void (*KernelFn)(const KHeader, const KArgs);
struct Variant {
  int nw, u1, ul;  // rollouts per CTA, dW tiles per warp of the first / last layer
  KernelFn fn[2][2];  // [loss and g = J^T r only][full checkpoints]
};
#define RL_VARIANT(NW, U1, UL)                                                                              \
  {NW, U1, UL, {{trx_rollout_kernel<NW, U1, UL, false, false>, trx_rollout_kernel<NW, U1, UL, false, true>}, \
                {trx_rollout_kernel<NW, U1, UL, true, false>, trx_rollout_kernel<NW, U1, UL, true, true>}}}
// smallest first: (2, 7) one-layer and narrow policies, (7, 13) the 41 -> 64 -> 83 policy of the benchmark (20 tiles
// = 80 of the 128 registers a thread of a 16-warp CTA can have), 4 rollouts per CTA when 8 regions do not fit the
// shared memory.  Wider policies have no re-rolled form here (TRX_ERR_UNSUPPORTED: use the tape objective).
const Variant kVariants[] = {RL_VARIANT(8, 2, 7), RL_VARIANT(8, 7, 13), RL_VARIANT(4, 7, 13)};

#define RL_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      err = std::string(#expr) + ": " + cudaGetErrorString(_e);                        \
      return TRX_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

}  // namespace

trx_status trxRolloutCreate(const char *options, uint64_t lanes, uint64_t max_device_bytes, int device,
                            bool scalar_rows, int checkpoint_mode, trx_rollout **out, std::string &err) {
  *out = nullptr;
  std::unique_ptr<trx_rollout, void (*)(trx_rollout *)> ro(new trx_rollout(), trxRolloutDestroy);
  try {
    dexsyn::ProblemOptions po = dexsyn::parseProblemOptions(options);
    if (po.outer != 1) {
      err = "re-rolled objective: outer batches are replicas of the flat tape only (use more rollouts)";
      return TRX_ERR_UNSUPPORTED;
    }
    ro->plan = makePlan(dexsyn::makeModel(po.model), dexsyn::EnvInfo(), po.frames);
  } catch (const std::exception &e) {
    err = e.what();
    return TRX_ERR_INVALID;
  }
  if (!ro->plan.error.empty()) {
    err = "re-rolled objective: " + ro->plan.error;
    return TRX_ERR_UNSUPPORTED;
  }
  if (lanes == 0) {
    err = "lanes must be positive";
    return TRX_ERR_INVALID;
  }
  const auto &I = ro->plan.i;
  if (I[H_CKPT] > 192) {
    err = "re-rolled objective: carried state exceeds the checkpoint prefetch registers (192 doubles)";
    return TRX_ERR_UNSUPPORTED;
  }
  if (I[H_NLAYERS] == 0 || I[H_NX] == 0) {
    err = "re-rolled objective: no policy, nothing to differentiate";
    return TRX_ERR_UNSUPPORTED;
  }
  std::memcpy(ro->hdr.v, I.data(), sizeof(ro->hdr.v));
  ro->device = device;
  ro->lanes = lanes;
  ro->scalar_rows = scalar_rows;
  ro->reg = ro->plan.d[I[H_D_CONST] + DC_W_REG];
  RL_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RL_CUDA(cudaGetDeviceProperties(&prop, device));

  // ---- shared-memory layout and kernel variant
  KArgs &a = ro->args;
  std::memset(&a, 0, sizeof(a));
  const int n_in = I[H_NIN], n_hid = I[H_NHID], n_out = I[H_NOUT], n_layers = I[H_NLAYERS];
  const int n1 = n_layers == 2 ? n_hid : n_out;
  a.ni = (int)I.size();
  a.nd = (int)ro->plan.d.size();
  a.nx = I[H_NX];
  a.frames = ro->plan.frames;
  a.tiles_j1 = (n1 + 7) / 8;
  a.tiles_1 = n_layers == 2 ? ((n_in + 7) / 8) * a.tiles_j1 : 0;
  a.tiles_jl = (n_out + 7) / 8;
  a.tiles_l = ((n_layers == 2 ? n_hid : n_in) + 7) / 8 * a.tiles_jl;
  const Variant *pick = nullptr;
  for (const Variant &v : kVariants) {
    int off = 0;
    auto take = [&](int n) {
      int o = off;
      off += padTo(n, 2);
      return o;
    };
    a.o_D = take(a.nd);
    a.o_I = take((a.ni + 1) / 2);
    a.s_W1 = fragStride(n1);
    a.r_W1 = padTo(n_in, 8);
    a.o_W1 = take(a.r_W1 * a.s_W1);
    a.o_B1 = take(a.s_W1);
    a.s_W2 = n_layers == 2 ? fragStride(n_out) : 0;
    a.r_W2 = n_layers == 2 ? padTo(n_hid, 8) : 0;
    a.o_W2 = take(a.r_W2 * a.s_W2);
    a.o_B2 = take(a.s_W2);
    a.s_X = fragStride(n_in);
    a.s_H = n_layers == 2 ? fragStride(n_hid) : 0;
    a.s_O = fragStride(n_out);
    a.o_X = take(8 * a.s_X);
    a.o_H = take(8 * a.s_H);
    a.o_A = take(8 * a.s_H);
    a.o_O = take(8 * a.s_O);
    a.o_GX = take(8 * a.s_X);
    a.o_GH = take(8 * a.s_H);
    a.o_GO = take(8 * a.s_O);
    a.o_tile = take(kHelperWarps * (v.u1 + v.ul));
    a.o_zero = take(8 * std::max(a.s_O, std::max(a.s_X, a.s_H)) + 8);
    a.o_red = take(8);
    a.o_region = off;
    off += v.nw * I[H_REGION];
    size_t bytes = (size_t)off * sizeof(double);
    if ((a.tiles_1 + kDwWarps - 1) / kDwWarps > v.u1 || (a.tiles_l + kDwWarps - 1) / kDwWarps > v.ul) continue;
    if (bytes > (size_t)prop.sharedMemPerBlockOptin) continue;
    if ((n_layers == 2 ? n_hid : 0) + n_out > 2 * kDwWarps * 32) continue;  // bias gradients: two per dW thread
    pick = &v;
    ro->smem_bytes = bytes;
    break;
  }
  if (!pick) {
    err = "re-rolled objective: policy layers of this size do not fit the kernel variants (shared memory / tiles per warp)";
    return TRX_ERR_UNSUPPORTED;
  }
  ro->nw = pick->nw;
  ro->u1 = pick->u1;
  ro->ul = pick->ul;
  for (int i = 0; i < 4; i++)
    RL_CUDA(cudaFuncSetAttribute(pick->fn[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ro->smem_bytes));

  // ---- device buffers; rollouts are walked in chunks when the checkpoints exceed the byte budget
  // what a checkpoint holds beyond the carried state: the more, the less the adjoint sweep
  // recomputes and the more crosses HBM.  TRX_ROLLOUT_CHECKPOINT = 0 | 1 | 2 overrides the default.
  a.ckpt_mode = checkpoint_mode;
  if (const char *w = getenv("TRX_ROLLOUT_CHECKPOINT")) a.ckpt_mode = std::max(0, std::min(2, atoi(w)));
  if (n_layers == 0) a.ckpt_mode = std::min(a.ckpt_mode, 1);  // the full restore / store is synchronised inside the policy's phases
  auto even = [](int n) { return (n + 1) & ~1; };
  a.ckpt_doubles = even(I[H_CKPT]);
  if (a.ckpt_mode >= 1) a.ckpt_doubles += 2 * even(n_hid) + even(n_out) + even(I[H_NQ]);
  if (a.ckpt_mode == 2) a.ckpt_doubles += even(I[H_NJ] * 7) + even(n_in) + I[H_NEE] * 12 + 4;
  a.ckpt_doubles = (a.ckpt_doubles + 15) / 16 * 16;  // whole 128-byte lines
  const uint64_t ck_bytes_per_rollout = (uint64_t)a.frames * a.ckpt_doubles * sizeof(double);
  uint64_t chunk = lanes;
  if (max_device_bytes && chunk * ck_bytes_per_rollout > max_device_bytes)
    chunk = std::max<uint64_t>(ro->nw, max_device_bytes / ck_bytes_per_rollout / ro->nw * ro->nw);
  ro->chunk = chunk;
  ro->n_chunks = (lanes + chunk - 1) / chunk;
  const uint64_t groups = (chunk + ro->nw - 1) / ro->nw;
  ro->grid = (int)std::min<uint64_t>(groups, (uint64_t)prop.multiProcessorCount);
  RL_CUDA(cudaMalloc(&ro->d_I, I.size() * sizeof(int32_t)));
  RL_CUDA(cudaMemcpy(ro->d_I, I.data(), I.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  RL_CUDA(cudaMalloc(&ro->d_D, ro->plan.d.size() * sizeof(double)));
  RL_CUDA(cudaMemcpy(ro->d_D, ro->plan.d.data(), ro->plan.d.size() * sizeof(double), cudaMemcpyHostToDevice));
  RL_CUDA(cudaMalloc(&ro->d_params, 3 * lanes * sizeof(double)));
  RL_CUDA(cudaMemset(ro->d_params, 0, 3 * lanes * sizeof(double)));
  RL_CUDA(cudaMalloc(&ro->d_ckpt, groups * ro->nw * ck_bytes_per_rollout));
  RL_CUDA(cudaMalloc(&ro->d_partial, (size_t)ro->grid * (1 + a.nx) * sizeof(double)));
  RL_CUDA(cudaMemset(ro->d_partial, 0, (size_t)ro->grid * (1 + a.nx) * sizeof(double)));
  a.I = ro->d_I;
  a.D = ro->d_D;
  a.params = ro->d_params;
  a.ckpt = ro->d_ckpt;
  a.partial = ro->d_partial;
  a.R = (long)lanes;
  *out = ro.release();
  return TRX_OK;
}

void trxRolloutDestroy(trx_rollout *ro) {
  if (!ro) return;
  cudaFree(ro->d_I);
  cudaFree(ro->d_D);
  cudaFree(ro->d_params);
  cudaFree(ro->d_ckpt);
  cudaFree(ro->d_partial);
  delete ro;
}

uint64_t trxRolloutInfo(const trx_rollout *ro, const char *key) {
  const auto &I = ro->plan.i;
  std::string k = key;
  if (k == "n_x") return I[H_NX];
  if (k == "parameter_elems") return 3 * ro->lanes;
  if (k == "chunk_lanes") return ro->chunk;
  if (k == "n_chunks") return ro->n_chunks;
  if (k == "device_bytes")
    return (uint64_t)((ro->chunk + ro->nw - 1) / ro->nw * ro->nw) * ro->plan.frames * ro->args.ckpt_doubles * 8 +
           (uint64_t)ro->grid * (1 + I[H_NX]) * 8 + 3 * ro->lanes * 8;
  if (k == "launches_per_evaluation") return ro->n_chunks + 1;
  if (k == "residual_elems") return (uint64_t)I[H_NSCALAR] + ((uint64_t)ro->plan.frames * I[H_ROWS_FRAME] + 6) * ro->lanes;
  if (k == "n_scalar_rows") return I[H_NSCALAR];
  if (k == "rows_per_frame") return I[H_ROWS_FRAME];
  if (k == "checkpoint_bytes_per_unit") return (uint64_t)ro->args.ckpt_doubles * 8;
  if (k == "carried_state_bytes_per_unit") return (uint64_t)I[H_CKPT] * 8;
  if (k == "checkpoint_mode") return (uint64_t)ro->args.ckpt_mode;
  if (k == "rollouts_per_cta") return ro->nw;
  if (k == "grid") return ro->grid;
  if (k == "shared_bytes") return ro->smem_bytes;
  if (k == "tiles_per_warp") return ro->u1 + ro->ul;
  if (k == "n_joints") return I[H_NJ];
  if (k == "n_levels") return I[H_NSUPER];
  if (k == "n_controlled") return I[H_NQ];
  if (k == "n_end_effectors") return I[H_NEE];
  if (k == "n_in") return I[H_NIN];
  if (k == "n_hidden") return I[H_NHID];
  if (k == "n_out") return I[H_NOUT];
  return ~0ull;
}

trx_status trxRolloutParameterize(trx_rollout *ro, const void *packed, uint64_t bytes, cudaStream_t stream,
                                  std::string &err) {
  if (bytes != 3 * ro->lanes * sizeof(double)) {
    err = "packed buffer size mismatch: got " + std::to_string(bytes) + " expected " +
          std::to_string(3 * ro->lanes * sizeof(double));
    return TRX_ERR_SIZE;
  }
  RL_CUDA(cudaSetDevice(ro->device));
  RL_CUDA(cudaMemcpyAsync(ro->d_params, packed, bytes, cudaMemcpyHostToDevice, stream));
  RL_CUDA(cudaStreamSynchronize(stream));
  return TRX_OK;
}

// d_out = {loss, gradient[n_x]} (want_grad) or {loss}; d_r: residual vector in the wide layout or null;
// d_seed: adjoint seeds in the same layout or null
trx_status trxRolloutEvaluate(trx_rollout *ro, const double *d_x, double *d_out, double *d_r, const double *d_seed,
                              bool want_grad, cudaStream_t stream, std::string &err) {
  RL_CUDA(cudaSetDevice(ro->device));
  KernelFn fn = nullptr;
  for (const Variant &v : kVariants)
    if (v.nw == ro->nw && v.u1 == ro->u1 && v.ul == ro->ul) fn = v.fn[(d_r || d_seed) ? 0 : 1][ro->args.ckpt_mode == 2 ? 1 : 0];
  KArgs a = ro->args;
  a.x = d_x;
  a.r_out = d_r;
  a.seed = d_seed;
  a.want_grad = want_grad ? 1 : 0;
  for (uint64_t ch = 0; ch < ro->n_chunks; ch++) {
    a.first = (long)(ch * ro->chunk);
    a.count = (long)std::min<uint64_t>(ro->chunk, ro->lanes - ch * ro->chunk);
    a.n_groups = (int)((a.count + ro->nw - 1) / ro->nw);
    a.accumulate = ch > 0 ? 1 : 0;
    // every CTA of the grid writes its partial slot, also when it has no group in this chunk
    fn<<<ro->grid, (ro->nw + kHelperWarps) * 32, ro->smem_bytes, stream>>>(ro->hdr, a);
    ro->launches++;
  }
  const int nx = ro->args.nx;
  trx_rollout_reduce_kernel<<<(nx + 255) / 256, 256, 0, stream>>>(ro->d_partial, ro->grid, nx, d_x, ro->reg,
                                                                    ro->scalar_rows ? 1 : 0, want_grad ? 1 : 0, d_seed,
                                                                    d_r, d_out);
  ro->launches++;
  RL_CUDA(cudaGetLastError());
  return TRX_OK;
}

uint64_t trxRolloutLaunches(const trx_rollout *ro) { return ro->launches; }

#ifdef TRX_ROLLOUT_PROFILE
extern "C" void trx_rollout_profile_dump() {
  long long h[64];
  cudaMemcpyFromSymbol(h, g_rollout_prof, sizeof(h));
  const char *names[64] = {};
  names[0] = "sim", names[1] = "chain", names[2] = "observe", names[3] = "barriers(fwd)", names[4] = "dense1", names[5] = "hidden";
  names[6] = "dense2", names[7] = "control", names[8] = "contact1", names[9] = "contact2", names[10] = "contact3";
  names[15] = "barrier B (body)", names[50] = "dW warp: last layer", names[51] = "dW warp: first layer", names[46] = "barrier C", names[47] = "barrier D (dW last)", names[48] = "barrier E (dW first)", names[49] = "barrier G (body adj)", names[11] = "contactsum", names[12] = "beginframe+ckpt", names[13] = "loadckpt";
  names[32] = "b.begin", names[33] = "b.contact3", names[34] = "b.contact2", names[35] = "b.contact1", names[36] = "b.control";
  names[37] = "b.barriers", names[38] = "b.dense2 gx", names[39] = "b.dW2+db2", names[40] = "b.hidden", names[41] = "b.dense1 gx";
  names[42] = "b.dW1+db1", names[43] = "b.observe", names[44] = "b.chain", names[45] = "b.sim";
  long long total = 0;
  for (int i = 0; i < 64; i++) total += h[i];
  for (int i = 0; i < 64; i++) {
    if (!h[i]) continue;
    const char *n = names[i];
    std::printf("%-2d %-22s %12lld cycles %5.1f%%\n", i, n ? n : "?", h[i], 100.0 * h[i] / total);
  }
}
#endif
