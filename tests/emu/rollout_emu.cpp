// rollout_emu.cpp -- TEST-ONLY host execution of the re-rolled rollout objective
// (csrc/rollout/rollout_frame.h): the same phase functions the sm_100a kernel runs, driven lane by
// lane on the CPU so that the hand-written forward and adjoint can be compared with the reference's
// goldens without a GPU.  The policy layers, which the kernel computes with DMMA tiles across the
// rollouts of a CTA, are plain loops here.  Never loaded by the package.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../robio_2021_dexopt_b200/csrc/assembly/dexsyn_options.h"
#include "../../robio_2021_dexopt_b200/csrc/rollout/rollout_frame.h"

using namespace rollout;

namespace {
std::string g_error;

struct Problem {
  Plan plan;
  int frames;
};

Problem makeProblem(const char *options) {
  dexsyn::ProblemOptions po = dexsyn::parseProblemOptions(options);
  if (po.outer != 1) throw std::runtime_error("outer batches are replicas of the flat tape only");
  Problem p;
  p.frames = po.frames;
  p.plan = makePlan(dexsyn::makeModel(po.model), dexsyn::EnvInfo(), po.frames);
  if (!p.plan.error.empty()) throw std::runtime_error(p.plan.error);
  return p;
}

// y[j] = sum_i x[i] W[i][j] + b[j]
void denseForward(const double *x, const double *W, const double *b, int n_in, int n_out, double *y) {
  for (int j = 0; j < n_out; j++) {
    double acc = 0;
    for (int i = 0; i < n_in; i++) acc += x[i] * W[i * n_out + j];
    y[j] = acc + b[j];
  }
}
// gx = gy W^T ; gW += x^T gy ; gb += gy
void denseReverse(const double *x, const double *gy, const double *W, int n_in, int n_out, double *gx, double *gW,
                  double *gb) {
  for (int i = 0; i < n_in; i++) {
    double acc = 0;
    for (int j = 0; j < n_out; j++) {
      acc += gy[j] * W[i * n_out + j];
      gW[i * n_out + j] += x[i] * gy[j];
    }
    gx[i] = acc;
  }
  for (int j = 0; j < n_out; j++) gb[j] += gy[j];
}
}  // namespace

#define LANES(stmt) for (int lane = 0; lane < 32; lane++) { stmt; }

// lane exchange of the warp-scan phases on 32-element arrays (the device uses warp shuffles)
struct HostWarp {
  template <class T> static void up(const T (&x)[32], int d, T (&o)[32]) {
    for (int i = 0; i < 32; i++) o[i] = i >= d ? x[i - d] : x[i];
  }
  template <class T> static void down(const T (&x)[32], int d, T (&o)[32]) {
    for (int i = 0; i < 32; i++) o[i] = i + d < 32 ? x[i + d] : x[i];
  }
};

extern "C" {

const char *rollout_emu_last_error() { return g_error.c_str(); }

int rollout_emu_info(const char *options, const char *key, long *value) {
  try {
    Problem p = makeProblem(options);
    const auto &I = p.plan.i;
    std::string k = key;
    if (k == "n_x") *value = I[H_NX];
    else if (k == "n_scalar") *value = I[H_NSCALAR];
    else if (k == "lane_rows") *value = (long)p.frames * I[H_ROWS_FRAME] + 6;
    else if (k == "rows_frame") *value = I[H_ROWS_FRAME];
    else if (k == "checkpoint_doubles") *value = I[H_CKPT];
    else if (k == "region_doubles") *value = I[H_REGION];
    else if (k == "n_joints") *value = I[H_NJ];
    else if (k == "n_levels") *value = I[H_NSUPER];
    else throw std::runtime_error("unknown key " + k);
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

// params: wide [3][lanes]; x: n_x; r_wide: n_scalar + lane_rows * lanes (or null); g: n_x (or null);
// seed: same layout as r_wide (or null: seeds are the residuals)
int rollout_emu_eval(const char *options, int lanes, const double *params, const double *x, double *r_wide,
                     double *loss_out, double *g, const double *seed_wide) {
  try {
    Problem p = makeProblem(options);
    const auto &I = p.plan.i;
    const auto &D = p.plan.d;
    const int T = p.frames, nq = I[H_NQ], n_in = I[H_NIN], n_hid = I[H_NHID], n_out = I[H_NOUT];
    const int n_layers = I[H_NLAYERS], L = I[H_NSUPER];
    const long nx = I[H_NX], n_scalar = I[H_NSCALAR];
    const double *W1 = x + I[H_W1], *b1 = x + I[H_B1], *W2 = x + I[H_W2], *b2 = x + I[H_B2];
    const double reg = D[I[H_D_CONST] + DC_W_REG];
    double loss_total = 0;
    std::vector<double> gsum(nx, 0.0);
    // lane-independent rows: weight decay (neural.h:159-171), counted once
    for (long i = 0; i < n_scalar; i++) {
      double v = x[i] * reg;
      loss_total += v * v;
      if (r_wide) r_wide[i] = v;
      double sd = seed_wide ? seed_wide[i] : v;
      gsum[i] += sd * reg;
    }
    for (int g0 = 0; g0 < lanes; g0++) {
      std::vector<double> region(I[H_REGION], 0.0), ckpt((size_t)T * I[H_CKPT]);
      std::vector<double> xs(n_in + 8), hs(n_hid + 8), as(n_hid + 8), os(n_out + 8), gxs(n_in + 8), ghs(n_hid + 8),
          gos(n_out + 8);
      Ctx c;
      c.H = I.data();
      c.I = I.data();
      c.D = D.data();
      c.s = region.data();
      c.x = xs.data(), c.h = hs.data(), c.a = as.data(), c.out = os.data();
      c.gx = gxs.data(), c.gh = ghs.data(), c.gout = gos.data();
      c.r_out = r_wide ? r_wide + n_scalar + g0 : nullptr;
      c.seed = nullptr;
      c.r_stride = lanes;
      c.frames = T;
      double lane_loss[32] = {0};
      double pr[3] = {params[0 * lanes + g0], params[1 * lanes + g0], params[2 * lanes + g0]};
      auto forwardKinematics = [&](const Ctx &cc) {
        for (int rd = 0; rd < I[H_NROUND]; rd++) fkRoundForward<32, HostWarp>(cc, rd, 0);
      };
      auto forwardFrame = [&](int f, const Ctx &cc, double *ll, bool store) {
        LANES(phaseSim(cc, lane));
        phaseBody(cc);
        forwardKinematics(cc);
        if (store) {
          LANES(phaseObserve<true>(cc, lane, f, ll[lane]));
        } else {
          LANES(phaseObserve<false>(cc, lane, f, ll[lane]));
        }
        if (n_layers == 2) {
          denseForward(cc.x, W1, b1, n_in, n_hid, cc.h);
          LANES(phaseHidden(cc, lane, f, ll[lane]));
          denseForward(cc.a, W2, b2, n_hid, n_out, cc.out);
        } else if (n_layers == 1) {
          denseForward(cc.x, W1, b1, n_in, n_out, cc.out);
        }
        if (n_layers) {
          LANES(phaseControl(cc, lane, f, ll[lane]));
          if (store) {
            LANES(phaseContact1<true>(cc, lane, f, ll[lane]));
          } else {
            LANES(phaseContact1<false>(cc, lane, f, ll[lane]));
          }
          LANES(phaseContact2(cc, lane, f));
          if (store) {
            LANES(phaseContact3(cc, lane, f, ll[lane]));
          }
        }
      };
      LANES(phaseInit(c, lane, pr));
      forwardKinematics(c);
      for (int f = 0; f < T; f++) {
        LANES(phaseBeginFrame(c, lane, f, &ckpt[(size_t)f * I[H_CKPT]]));
        forwardFrame(f, c, lane_loss, true);
      }
      LANES(phaseTerminal(c, lane, lane_loss[lane]));
      for (int lane = 0; lane < 32; lane++) loss_total += lane_loss[lane];
      if (!g) continue;
      // ---- adjoint sweep
      Ctx cb = c;
      cb.r_out = nullptr;
      cb.seed = seed_wide ? seed_wide + n_scalar + g0 : nullptr;
      double dummy[32] = {0};
      LANES(bphaseInit(cb, lane));
      for (int f = T - 1; f >= 0; f--) {
        const double *ck = &ckpt[(size_t)f * I[H_CKPT]];
        for (int lane = 0; lane < 32; lane++) {
          double pre[8];
          for (int i = 0; i < 8; i++) pre[i] = lane + 32 * i < I[H_CKPT] ? ck[lane + 32 * i] : 0.0;
          phaseLoadCheckpoint(cb, lane, pre);
        }
        forwardFrame(f, cb, dummy, false);
        LANES(bphaseBegin(cb, lane, f));
        if (f == T - 1 || f == 0) LANES(bphaseBegin2(cb, lane, f));
        if (n_layers) {
          LANES(bphaseContact3(cb, lane, f));
          LANES(bphaseContact2(cb, lane, f));
          LANES(bphaseContact1(cb, lane, f));
          LANES(bphaseControl(cb, lane, f));
          if (n_layers == 2) {
            denseReverse(cb.a, cb.gout, W2, n_hid, n_out, cb.gh, &gsum[I[H_W2]], &gsum[I[H_B2]]);
            LANES(bphaseHidden(cb, lane, f));
            denseReverse(cb.x, cb.gh, W1, n_in, n_hid, cb.gx, &gsum[I[H_W1]], &gsum[I[H_B1]]);
          } else {
            denseReverse(cb.x, cb.gout, W1, n_in, n_out, cb.gx, &gsum[I[H_W1]], &gsum[I[H_B1]]);
          }
        }
        LANES(bphaseObserve(cb, lane, f));
        LANES(bphaseObjectGather(cb, lane));
        for (int rd = I[H_NROUND] - 1; rd >= 0; rd--) fkRoundReverse<32, HostWarp>(cb, rd, 0);
        LANES(fkClearStatics(cb, lane));
        LANES(bphaseSim(cb, lane));
        bphaseBody(cb);
      }
    }
    *loss_out = loss_total;
    if (g) std::memcpy(g, gsum.data(), nx * sizeof(double));
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

}  // extern "C"
