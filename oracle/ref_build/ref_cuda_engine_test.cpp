// ref_cuda_engine_test.cpp -- the drop-in check: the SAME reference code path (record through Var<>,
// buildGradients, Executable::run/execute) driven once through the reference's SimpleEngine and once
// through tractor::CudaEngine (integration/cuda_engine.h -> include/trx_b200.h -> libtrx_b200.so),
// compared element by element.  Runs on the GPU box (tests/test_cuda_engine.py).  TEST INFRASTRUCTURE.
#include "ref_backend.h"
#include <tractor/core/gradients.h>
#include "../../integration/cuda_engine.h"
#include "../../robio_2021_dexopt_b200/csrc/assembly/dexsyn_assemble.h"

using namespace refc;
typedef tractor::Batch8d Lane;

static double maxRelErr(const std::vector<double> &a, const std::vector<double> &b) {
  if (a.size() != b.size()) return 1e300;
  double scale = 1.0, err = 0.0;
  for (double v : b) scale = std::max(scale, std::fabs(v));
  for (size_t i = 0; i < a.size(); i++) err = std::max(err, std::fabs(a[i] - b[i]) / scale);
  return err;
}

int main(int argc, char **argv) {
  try {
    int frames = argc > 1 ? std::atoi(argv[1]) : 2;
    dexsyn::ModelOptions mo;
    mo.policy_hidden = 4;
    mo.extra_fixed = 2;
    dexsyn::Model model = dexsyn::makeModel(mo);
    size_t np = dexsyn::Assembler<RefBackend<Lane>>::parameterCount(model);
    dexsyn::Rng rng(11);
    std::vector<double> x(np);
    for (auto &v : x) v = 0.1 * rng.normal();
    GradientSet gs;
    size_t n_params = 0;
    {
      MuteCout mute;
      gs.prog.record([&]() {
        RefBackend<Lane> b;
        dexsyn::Assembler<RefBackend<Lane>> as(b, model);
        as.rollout(frames, x);
        n_params = b.params.size();
      });
      gs.build();
    }
    std::vector<double> params(n_params * 8);
    for (size_t i = 0; i < params.size(); i++) params[i] = (i / 8 < 2) ? rng.uniform(-0.01, 0.01) : rng.uniform(0, 6.28);

    auto runAll = [&](const std::shared_ptr<tractor::Engine> &engine, std::vector<double> &r, std::vector<double> &g,
                      std::vector<double> &h) {
      auto x_prog = engine->compile(gs.prog);
      auto x_prep = engine->compile(gs.prep);
      auto x_bprop = engine->compile(gs.bprop);
      auto x_hprop = engine->compile(gs.hprop);
      auto memory = engine->createMemory();
      x_prog->parameterVector(params, memory);
      x_prog->run(x, memory, r);      // solvers/gd.h:50
      x_prep->execute(memory);        // :57
      x_bprop->run(r, memory, g);     // :62
      std::vector<double> dx(x.size());
      for (size_t i = 0; i < dx.size(); i++) dx[i] = 0.01 * std::sin(0.37 * (double)i);
      x_hprop->run(dx, memory, h);    // core/eigen.h:73-77 (Gauss-Newton product J^T J dx)
    };
    std::vector<double> r0, g0, h0, r1, g1, h1;
    runAll(std::make_shared<tractor::SimpleEngine>(), r0, g0, h0);
    runAll(std::make_shared<tractor::CudaEngine>(0), r1, g1, h1);
    // the other executables SolverBase compiles for every solver (solvers/base.h:112-137): accu, and the seven
    // programs of buildConstraints -- for a tape without constraint operators project is a move per variable, the
    // barrier programs zeros, the penalty programs empty
    tractor::Program proj, b_init, b_step, b_diag, p_init, p_step, p_diag;
    tractor::buildConstraints(gs.prog, gs.fprop, gs.bprop, gs.hprop, tractor::TypeInfo::get<double>(), &proj, &b_init,
                              &b_step, &b_diag, &p_init, &p_step, &p_diag);
    auto runRest = [&](const std::shared_ptr<tractor::Engine> &engine, std::vector<double> &out) {
      auto memory = engine->createMemory();
      auto x_prog = engine->compile(gs.prog);
      auto x_accu = engine->compile(gs.accu);
      auto x_proj = engine->compile(proj), x_bi = engine->compile(b_init), x_bs = engine->compile(b_step),
           x_bd = engine->compile(b_diag), x_pi = engine->compile(p_init), x_ps = engine->compile(p_step),
           x_pd = engine->compile(p_diag);
      std::vector<double> r, y, step(x.size()), xd(2 * x.size());
      for (size_t i = 0; i < step.size(); i++) step[i] = 0.02 * std::cos(0.11 * (double)i);
      x_prog->parameterVector(params, memory);
      x_prog->run(x, memory, r);
      out.clear();
      // accumulate: x (+) step   (solvers/base.h:158-165: inputs x then the step)
      for (size_t i = 0; i < x.size(); i++) xd[i] = x[i], xd[x.size() + i] = step[i];
      x_accu->run(xd, memory, y);
      out.insert(out.end(), y.begin(), y.end());
      x_proj->parameterVector(std::vector<double>(1, 0.1), memory);
      x_proj->run(step, memory, y);
      out.insert(out.end(), y.begin(), y.end());
      x_bi->run(step, memory, y);
      out.insert(out.end(), y.begin(), y.end());
      x_bs->run(step, memory, y);
      out.insert(out.end(), y.begin(), y.end());
      x_bd->execute(memory);
      x_bd->outputVector(memory, y);
      out.insert(out.end(), y.begin(), y.end());
      x_pi->execute(memory);
      x_ps->execute(memory);
      x_pd->execute(memory);
    };
    std::vector<double> rest0, rest1;
    runRest(std::make_shared<tractor::SimpleEngine>(), rest0);
    runRest(std::make_shared<tractor::CudaEngine>(0), rest1);
    const double erest = maxRelErr(rest1, rest0);
    std::printf("{\"solverbase_executables\": 13, \"values\": %zu, \"err_rest\": %.3g}\n", rest0.size(), erest);
    if (!(erest < 1e-12)) return 3;
    double er = maxRelErr(r1, r0), eg = maxRelErr(g1, g0), eh = maxRelErr(h1, h0);
    std::printf("{\"residuals\": %zu, \"gradient\": %zu, \"err_r\": %.3g, \"err_g\": %.3g, \"err_h\": %.3g}\n", r0.size(),
                g0.size(), er, eg, eh);
    return (er < 1e-9 && eg < 1e-9 && eh < 1e-9) ? 0 : 2;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "ref_cuda_engine_test: %s\n", e.what());
    return 1;
  }
}
