// ref_tool.cpp -- drives the REFERENCE's own tape library (libtractor_ref.so,
// built from /root/reference by this directory's Makefile) to produce the
// fixtures and CPU baselines every parity claim is pinned to.
//
//   ref_tool registry <out.txt>
//       dump Operator::all(): name, argument sizes/directions/lane types
//   ref_tool optests <out.trxb> [suffix=8d|d] [seed=N]
//       one single-instruction program per registered compute op plus its
//       buildGradients() derivatives, random inputs, and SimpleEngine outputs
//       (same construction as tractor/src/test/test.cpp:565-588)
//   ref_tool record <out.trxb> problem=dexsyn|fkchain kind=... frames=T tiles=K ...
//       record a synthetic problem through the reference Var<> API, derive
//       prep/fprop/bprop/hprop/accu with the reference's buildGradients(), run
//       the reference SimpleEngine on K lane tiles and store residuals,
//       gradient, tangent and Gauss-Newton products as golden arrays
//   ref_tool bench <in.trxb> threads=N seconds=S [engine=simple|loop]
//       time objective+gradient (x_prog, x_prep, x_bprop) on N host threads,
//       each with its own Memory (engine.h:44-48 are const)
//
// TEST INFRASTRUCTURE ONLY.
#include <atomic>
#include <cstring>
#include <deque>
#include <fstream>
#include <mutex>
#include <thread>

#include "ref_backend.h"
#include <tractor/core/constraints.h>
#include <tractor/core/gradients.h>
#include "../../robio_2021_dexopt_b200/csrc/assembly/dexsyn_assemble.h"

#include "../../robio_2021_dexopt_b200/csrc/collision/trx_shape_table.h"

using namespace refc;
typedef tractor::Batch8d Lane;

// ---- collision problems: the model's shapes as stand-in reference objects whose queries go through the hook
// (shim/tractor_collision/collision_decl.h).  The hook is answered by the geometric query under test
// (csrc/collision/trx_collide.h, compiled here for the host) -- parity of that arithmetic against Bullet stays
// unpinned; everything around it on the tape is the reference's own code.
static trxcol::ShapeTable g_shapes;
static std::deque<tractor::CollisionShape<double>> g_shape_objects;
static std::deque<tractor::ShapeCollisionPair<double>> g_pair_objects;
static void hookProject(uint64_t shape, const double *p, double *n, double *d) {
  trxvm::V3 nrm{0, 0, 0};
  double dist = 0;
  trxcol::projectPoint(g_shapes.heap.data(), shape, trxvm::V3{p[0], p[1], p[2]}, nrm, dist);
  n[0] = nrm.x, n[1] = nrm.y, n[2] = nrm.z, *d = dist;
}
static void hookPair(uint64_t pair, const double *pa, const double *pb, double *a, double *b, double *n) {
  trxvm::PoseV A, B;
  A.t = {pa[0], pa[1], pa[2]}, A.r = {pa[3], pa[4], pa[5], pa[6]};
  B.t = {pb[0], pb[1], pb[2]}, B.r = {pb[3], pb[4], pb[5], pb[6]};
  trxcol::Result r;
  trxcol::collideShapes(g_shapes.heap.data(), pair, A, B, r);
  a[0] = r.a.x, a[1] = r.a.y, a[2] = r.a.z, b[0] = r.b.x, b[1] = r.b.y, b[2] = r.b.z, n[0] = r.n.x, n[1] = r.n.y, n[2] = r.n.z;
}
// the same shapes and pairs, in the same order, as csrc/trx_builder.cpp registerShapes()
static void registerShapes(dexsyn::Model &model) {
  if (!model.wants_collision_pairs && !model.friction_cones) return;
  tractor::collisionQueryHook().project = hookProject;
  tractor::collisionQueryHook().pair = hookPair;
  auto reg = [](dexsyn::Shape &s) {
    trx_shape_desc d;
    std::memset(&d, 0, sizeof(d));
    std::vector<double> planes;
    if (s.kind == 1) {
      d.kind = TRX_SHAPE_CYLINDER, d.radius = s.radius, d.length = s.length;
    } else {
      d.kind = TRX_SHAPE_POLYHEDRON;
      for (auto &pl : s.planes) planes.insert(planes.end(), {pl.n[0], pl.n[1], pl.n[2], pl.d});
      d.n_points = (uint32_t)(s.points.size() / 3), d.n_planes = (uint32_t)s.planes.size();
      d.points = s.points.data(), d.planes = planes.data();
    }
    uint64_t h = 0;
    bool unsupported = false;
    std::string err = g_shapes.addShape(d, h, unsupported);
    if (!err.empty()) throw std::runtime_error(err);
    g_shape_objects.emplace_back();
    g_shape_objects.back().handle = h;
    s.handle = (uint64_t)&g_shape_objects.back();  // the tape carries the object's address (dexlearn.h:329)
    return h;
  };
  const uint64_t h_obj = reg(model.object_shape);
  std::vector<uint64_t> h_ee;
  for (auto &s : model.ee_shapes) h_ee.push_back(reg(s));
  if (!model.wants_collision_pairs) return;
  auto pair = [&](int ea, uint64_t ha, int link_b, uint64_t hb) {
    dexsyn::CollisionPair cp;
    cp.link_a = model.end_effectors[ea];
    cp.link_b = link_b;
    uint64_t h = 0;
    std::string err = g_shapes.addPair(ha, hb, h);
    if (!err.empty()) throw std::runtime_error(err);
    g_pair_objects.emplace_back();
    g_pair_objects.back().handle = h;
    cp.handle = (uint64_t)&g_pair_objects.back();
    model.collision_pairs.push_back(cp);
  };
  const int n_ee = (int)model.end_effectors.size();
  for (int e = 0; e < n_ee - (model.floor_pair ? 0 : 1); e++) pair(e, h_ee[e], model.object_joint, h_obj);
  for (int e = 0; e + 2 < n_ee; e++) pair(e, h_ee[e], model.end_effectors[e + 1], h_ee[e + 1]);
}

static std::map<std::string, std::string> parseArgs(int argc, char **argv, int first) {
  std::map<std::string, std::string> kv;
  for (int i = first; i < argc; i++) {
    std::string a = argv[i];
    auto eq = a.find('=');
    if (eq == std::string::npos) throw std::runtime_error("expected key=value: " + a);
    kv[a.substr(0, eq)] = a.substr(eq + 1);
  }
  return kv;
}
static std::string getS(const std::map<std::string, std::string> &kv, const std::string &k, const std::string &d) {
  auto it = kv.find(k);
  return it == kv.end() ? d : it->second;
}
static long getI(const std::map<std::string, std::string> &kv, const std::string &k, long d) {
  auto it = kv.find(k);
  return it == kv.end() ? d : std::atol(it->second.c_str());
}
static double getD(const std::map<std::string, std::string> &kv, const std::string &k, double d) {
  auto it = kv.find(k);
  return it == kv.end() ? d : std::atof(it->second.c_str());
}

static const char *modeName(const tractor::Operator *op) {
  if (op->isMode<tractor::compute>()) return "compute";
  if (op->isMode<tractor::prepare>()) return "prepare";
  if (op->isMode<tractor::forward>()) return "forward";
  if (op->isMode<tractor::reverse>()) return "reverse";
  return "other";
}

// ------------------------------------------------------------------ registry

static int cmdRegistry(const std::string &path) {
  auto ops = tractor::Operator::all();
  std::sort(ops.begin(), ops.end(),
            [](const tractor::Operator *a, const tractor::Operator *b) { return a->name() < b->name(); });
  std::ofstream f(path);
  for (auto *op : ops) {
    const std::string &n = op->name();
    bool d = n.size() > 2 && n.substr(n.size() - 2) == "_d";
    bool d8 = n.size() > 3 && n.substr(n.size() - 3) == "_8d";
    if (!d && !d8) continue;
    f << n << " " << modeName(op) << " " << op->argumentCount();
    for (auto &arg : op->arguments()) {
      uint8_t lanes, elem;
      classify(arg.typeInfo(), lanes, elem);
      f << " " << arg.size() << ":" << (arg.isInput() ? "i" : "o") << ":" << (int)lanes << ":" << (int)elem;
    }
    f << "\n";
  }
  return 0;
}

// ------------------------------------------------------------------ per-op tests

static bool endsWith(const std::string &s, const std::string &suffix) {
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

// Every registered operator whose name contains "constraint" (core/constraints.h, geometry/ops.h:998-1130), in every
// mode -- compute, prepare, forward, reverse, project, barrier_init / _step / _diagonal, penalty_* -- as its own
// single-instruction program: random inputs, the reference SimpleEngine's outputs.
static int cmdConstraintTests(const std::string &path, const std::map<std::string, std::string> &kv) {
  std::string suffix = "_" + getS(kv, "suffix", "8d");
  dexsyn::Rng rng((uint64_t)getI(kv, "seed", 11));
  auto ops = tractor::Operator::all();
  std::sort(ops.begin(), ops.end(),
            [](const tractor::Operator *a, const tractor::Operator *b) { return a->name() < b->name(); });
  trxfile::Bundle bundle;
  std::string listing;
  for (auto *op : ops) {
    const std::string &n = op->name();
    if (n.find("constraint") == std::string::npos || !endsWith(n, suffix)) continue;
    bool has_output = false;
    for (auto &arg : op->arguments()) has_output |= !arg.isInput();
    if (!has_output) continue;  // prepare rules of constraints save nothing
    tractor::Program prog;
    std::vector<tractor::Program::Instruction> insts;
    // every argument in its own 512-byte cell: the reference's quaternion / vector3 barrier rules index six
    // components of three-component arguments (geometry/ops.h:1076-1081,1106-1125); the padding keeps those stray
    // accesses away from the neighbouring arguments, so the in-range components are well defined
    size_t addr = 512;
    insts.push_back((uintptr_t)op);
    size_t in_off = 0, out_off = 0;
    for (auto &arg : op->arguments()) {
      insts.push_back(addr);
      if (arg.isInput()) {
        prog.addInput(tractor::Program::Input(arg.typeInfo(), addr, in_off, 0));
        in_off += arg.size();
      } else {
        prog.addOutput(tractor::Program::Output(arg.typeInfo(), addr, out_off, 0));
        out_off += arg.size();
      }
      addr += 512;
    }
    prog.updateMemorySize(addr + 512);
    prog.setInstructions(insts);
    size_t nx = portElements(prog.inputs());
    std::vector<double> x(nx), y;
    for (auto &v : x) v = rng.uniform(-1, 1);
    auto engine = std::make_shared<tractor::SimpleEngine>();
    auto memory = engine->createMemory();
    auto x_prog = engine->compile(prog);
    x_prog->run(x, memory, y);
    std::string prefix = "op/" + n + "/";
    bundle.programs.emplace_back(prefix + "prog", convertProgram(prog));
    bundle.arrays.emplace_back(prefix + "x", x);
    bundle.arrays.emplace_back(prefix + "y", y);
    listing += n + "\n";
  }
  bundle.texts.emplace_back("ops", listing);
  trxfile::writeBundle(bundle, path);
  return 0;
}

// A small bounded least-squares tape whose variables carry constraints, as robot/jointtypes.h:105-119 records them
// for joint variables: goals (w_i - target_i per lane), goal(range_constraint(w, lo, hi)) /
// goal(trust_region_constraint(w, tr)) / goal(range_trust_region_constraint(w, lo, hi, tr)).  The reference's
// buildGradients + buildConstraints (gradients.cpp:22-407,764-997) derive the programs SolverBase compiles
// (solvers/base.h:117-137); the reference engine runs them the way solvers/ip.h:212-246,287-293,475-478 does.
static int cmdConstrained(const std::string &path, const std::map<std::string, std::string> &kv) {
  const int n = (int)getI(kv, "variables", 12);
  dexsyn::Rng rng((uint64_t)getI(kv, "seed", 5));
  GradientSet gs;
  std::deque<tractor::Var<double>, tractor::AlignedStdAlloc<tractor::Var<double>>> vars;
  std::deque<tractor::Var<Lane>, tractor::AlignedStdAlloc<tractor::Var<Lane>>> params;
  std::vector<double> x0(n);
  {
    MuteCout mute;
    gs.prog.record([&]() {
      for (int i = 0; i < n; i++) {
        x0[i] = rng.uniform(-0.5, 0.5);
        vars.emplace_back(x0[i]);
        tractor::variable(vars.back());
        params.emplace_back(Lane(0));
        tractor::parameter(params.back());
        tractor::Var<Lane> wi;
        tractor::batch(vars.back(), wi);
        tractor::goal(tractor::sub(wi, params.back()));
        const double lo = x0[i] - rng.uniform(0.05, 0.6), hi = x0[i] + rng.uniform(0.05, 0.6), tr = rng.uniform(0.05, 0.4);
        if (i % 4 == 0) tractor::goal(tractor::range_constraint(vars.back(), tractor::Var<double>(lo), tractor::Var<double>(hi)));
        else if (i % 4 == 1) tractor::goal(tractor::trust_region_constraint(vars.back(), tractor::Var<double>(tr)));
        else if (i % 4 == 2)
          tractor::goal(tractor::range_trust_region_constraint(vars.back(), tractor::Var<double>(lo), tractor::Var<double>(hi),
                                                               tractor::Var<double>(tr)));
        // (i % 4 == 3: an unconstrained variable -- buildConstraints passes it through: move / zero)
      }
    });
  }
  gs.build();
  tractor::Program proj, b_init, b_step, b_diag, p_init, p_step, p_diag;
  tractor::buildConstraints(gs.prog, gs.fprop, gs.bprop, gs.hprop, tractor::TypeInfo::get<double>(), &proj, &b_init, &b_step,
                            &b_diag, &p_init, &p_step, &p_diag);
  trxfile::Bundle bundle;
  gs.addTo(bundle);
  bundle.programs.emplace_back("project", convertProgram(proj));
  bundle.programs.emplace_back("barrier_init", convertProgram(b_init));
  bundle.programs.emplace_back("barrier_step", convertProgram(b_step));
  bundle.programs.emplace_back("barrier_diagonal", convertProgram(b_diag));
  auto engine = std::make_shared<tractor::SimpleEngine>();
  auto memory = engine->createMemory();
  auto x_prog = engine->compile(gs.prog);
  auto x_proj = engine->compile(proj), x_bi = engine->compile(b_init), x_bs = engine->compile(b_step),
       x_bd = engine->compile(b_diag);
  std::vector<double> par(portElements(gs.prog.parameters())), r, step(n), step2(n), padding(1, 0.05), y;
  for (auto &v : par) v = rng.uniform(-0.5, 0.5);
  for (auto &v : step) v = rng.uniform(-0.7, 0.7);  // some steps leave their ranges
  for (auto &v : step2) v = rng.uniform(-0.2, 0.2);
  x_prog->parameterVector(par, memory);
  x_prog->run(x0, memory, r);
  bundle.arrays.emplace_back("x", x0);
  bundle.arrays.emplace_back("tile0/params", par);
  bundle.arrays.emplace_back("tile0/r", r);
  bundle.arrays.emplace_back("step", step);
  bundle.arrays.emplace_back("step2", step2);
  bundle.arrays.emplace_back("padding", padding);
  x_proj->parameterVector(padding, memory);
  x_proj->run(step, memory, y);
  bundle.arrays.emplace_back("project/y", y);
  x_bi->run(step2, memory, y);
  bundle.arrays.emplace_back("barrier_init/y", y);
  x_bs->run(step, memory, y);
  bundle.arrays.emplace_back("barrier_step/y", y);
  x_bd->execute(memory);
  x_bd->outputVector(memory, y);
  bundle.arrays.emplace_back("barrier_diagonal/y", y);
  trxfile::writeBundle(bundle, path);
  return 0;
}

static int cmdOpTests(const std::string &path, const std::map<std::string, std::string> &kv) {
  std::string suffix = "_" + getS(kv, "suffix", "8d");
  dexsyn::Rng rng((uint64_t)getI(kv, "seed", 7));
  auto ops = tractor::Operator::all();
  std::sort(ops.begin(), ops.end(),
            [](const tractor::Operator *a, const tractor::Operator *b) { return a->name() < b->name(); });
  trxfile::Bundle bundle;
  std::string listing;
  for (auto *op : ops) {
    if (!op->isMode<tractor::compute>()) continue;
    if (!endsWith(op->name(), suffix)) continue;
    bool ok = true;
    for (auto &arg : op->arguments()) {
      uint8_t lanes, elem;
      classify(arg.typeInfo(), lanes, elem);
      if (elem != trxfile::ELEM_F64) ok = false;  // uint64 shape handles: host-callback ops
    }
    if (!ok) continue;
    if (op->name().find("constraint") != std::string::npos) continue;  // out of scope
    if (op->name().find("random") != std::string::npos || op->name().find("dropout") != std::string::npos)
      continue;  // in-tape RNG: not reproducible even reference-vs-reference

    // single-instruction program, arguments laid out back to back (test.cpp:565-588);
    // start at 64 so that every Batch argument is 32-byte aligned
    GradientSet gs;
    std::vector<tractor::Program::Instruction> insts;
    tractor::Allocator alloc(64);
    insts.push_back((uintptr_t)op);
    size_t in_off = 0, out_off = 0;
    for (auto &arg : op->arguments()) {
      size_t addr = alloc.alloc(arg.typeInfo());
      insts.push_back(addr);
      if (arg.isInput()) {
        gs.prog.addInput(tractor::Program::Input(arg.typeInfo(), addr, in_off, 0));
        in_off += arg.size();
      } else {
        gs.prog.addOutput(tractor::Program::Output(arg.typeInfo(), addr, out_off, 0));
        out_off += arg.size();
      }
    }
    gs.prog.updateMemorySize(alloc.top());
    gs.prog.setInstructions(insts);
    bool has_grad = op->tryFindVariant<tractor::forward>() && op->tryFindVariant<tractor::reverse>();
    std::string prefix = "op/" + op->name() + "/";
    size_t nx = portElements(gs.prog.inputs()), ny = portElements(gs.prog.outputs());
    std::vector<double> x(nx), y, gy(ny), gx, dx(nx), dy;
    const std::string &n = op->name();
    for (auto &v : x) v = rng.uniform(-1, 1);
    if (n.rfind("log_", 0) == 0 || n.rfind("sqrt_", 0) == 0)
      for (auto &v : x) v = std::fabs(v) + 0.1;
    if (n.rfind("div_", 0) == 0)
      for (size_t i = nx / 2; i < nx; i++) x[i] = (x[i] < 0 ? -1 : 1) * (std::fabs(x[i]) + 0.25);

    auto engine = std::make_shared<tractor::SimpleEngine>();
    auto memory = engine->createMemory();
    auto x_prog = engine->compile(gs.prog);
    x_prog->run(x, memory, y);
    bundle.programs.emplace_back(prefix + "prog", convertProgram(gs.prog));
    bundle.arrays.emplace_back(prefix + "x", x);
    bundle.arrays.emplace_back(prefix + "y", y);
    if (has_grad) {
      tractor::buildGradients(gs.prog, gs.prep, &gs.fprop, &gs.bprop);
      // tangent-space sizes (Pose -> Twist, Quaternion -> Vector3 for scalar double; the
      // gradient-type map has no Batch entries, type.h:55-58, so *_8d ports keep their size)
      gy.assign(portElements(gs.bprop.inputs()), 0.0);
      dx.assign(portElements(gs.fprop.inputs()), 0.0);
      for (auto &v : gy) v = rng.uniform(-1, 1);
      for (auto &v : dx) v = rng.uniform(-1, 1);
      auto x_prep = engine->compile(gs.prep);
      auto x_fprop = engine->compile(gs.fprop);
      auto x_bprop = engine->compile(gs.bprop);
      x_prep->execute(memory);
      x_bprop->run(gy, memory, gx);
      // fresh memory for the tangent sweep, as test.cpp:604-616 does
      memory = engine->createMemory();
      x_prog->run(x, memory, y);
      x_prep->execute(memory);
      x_fprop->run(dx, memory, dy);
      bundle.programs.emplace_back(prefix + "prep", convertProgram(gs.prep));
      bundle.programs.emplace_back(prefix + "fprop", convertProgram(gs.fprop));
      bundle.programs.emplace_back(prefix + "bprop", convertProgram(gs.bprop));
      bundle.arrays.emplace_back(prefix + "gy", gy);
      bundle.arrays.emplace_back(prefix + "gx", gx);
      bundle.arrays.emplace_back(prefix + "dx", dx);
      bundle.arrays.emplace_back(prefix + "dy", dy);
    }
    listing += op->name() + (has_grad ? " grad\n" : " nograd\n");
  }
  bundle.texts.emplace_back("ops", listing);
  trxfile::writeBundle(bundle, path);
  std::cerr << "wrote " << path << "\n" << listing;
  return 0;
}

// ------------------------------------------------------------------ problems

// Synthetic serial FK chain with pose-residual style goals: the smallest tape
// that exercises mul_fg_pose / fg_pose_angle_axis_pose and their adjoints.
static void recordFkChain(GradientSet &gs, int joints, uint64_t seed, size_t &n_x, size_t &n_param_lane) {
  dexsyn::ModelOptions mo;
  mo.kind = "chain";
  mo.chain_joints = joints;
  mo.chain_contacts = 3;
  mo.extra_fixed = 0;
  mo.policy_hidden = -1;
  mo.seed = seed;
  dexsyn::Model model = dexsyn::makeModel(mo);
  n_x = joints;
  n_param_lane = 3;
  MuteCout mute;
  gs.prog.record([&]() {
    typedef RefBackend<Lane> B;
    B b;
    // joint angles: shared scalar variables broadcast to the lanes, plus a per-lane offset parameter
    std::vector<B::S> q;
    B::S off = b.param();
    for (int i = 0; i < joints; i++) {
      B::W w = b.weight(0.1 * (i % 5) - 0.2);
      B::S qi;
      tractor::batch(w, qi);
      q.push_back(b.add(qi, off));
    }
    B::S py = b.param(), pz = b.param();
    B::P pose = b.cp(dexsyn::Joint().origin);
    int k = 0;
    for (size_t j = 0; j < model.joints.size(); j++) {
      auto &jt = model.joints[j];
      if (jt.type != dexsyn::JOINT_REVOLUTE) continue;
      pose = b.prevolute(b.pmul(pose, b.cp(jt.origin)), q[k], b.cv(jt.axis[0], jt.axis[1], jt.axis[2]));
      if (k % 8 == 7 || k == joints - 1) {
        B::V3 target = b.pack(b.cs(0.2), py, pz);
        b.goalV(b.vsub(b.pmulv(pose, b.cv(0.01, 0.02, 0.03)), target));
        b.goalV(b.qresidual(b.porientation(pose)));
      }
      k++;
    }
  });
  gs.build();
}

static dexsyn::Model g_model;

static void recordDexSyn(GradientSet &gs, const std::map<std::string, std::string> &kv, std::vector<double> &x0,
                         size_t &n_param_lane) {
  dexsyn::ModelOptions mo;
  mo.kind = getS(kv, "kind", "handarm");
  mo.chain_joints = (int)getI(kv, "joints", 30);
  mo.chain_contacts = (int)getI(kv, "contacts", 16);
  mo.extra_fixed = (int)getI(kv, "extra_fixed", 12);
  mo.policy_hidden = (int)getI(kv, "hidden", 64);
  mo.dropout = getI(kv, "dropout", 0) != 0;
  mo.collision = getI(kv, "collision", 0) != 0;
  mo.friction = getI(kv, "friction", 0) != 0;
  mo.floor_pair = getI(kv, "floor_pair", 1) != 0;
  mo.seed = (uint64_t)getI(kv, "seed", 20211227);
  int frames = (int)getI(kv, "frames", 4);
  int outer = (int)getI(kv, "outer", 1);
  double wstd = getD(kv, "wstd", 0.1);
  g_model = dexsyn::makeModel(mo);
  registerShapes(g_model);
  size_t np = dexsyn::Assembler<RefBackend<Lane>>::parameterCount(g_model);
  dexsyn::Rng rng(mo.seed + 1);
  x0.resize(np);
  for (auto &v : x0) v = wstd * rng.normal();
  MuteCout mute;
  size_t nparam = 0;
  gs.prog.record([&]() {
    RefBackend<Lane> b;
    dexsyn::Assembler<RefBackend<Lane>> as(b, g_model);
    // the reference's "outer batch" records the rollout again onto the same tape
    // with the same weights (dexlearn.h:440-453)
    for (int o = 0; o < outer; o++) as.rollout(frames, x0);
    nparam = b.params.size();
  });
  n_param_lane = nparam;
  gs.build();
}

// Runs K lane tiles through the reference engine and stores the golden arrays.
static void goldenRuns(const GradientSet &gs, trxfile::Bundle &bundle, const std::vector<double> &x, size_t n_param_lane,
                       int tiles, uint64_t seed, const std::string &param_kind, int gd_steps = 0,
                       double learning_rate = 0.01, bool gd_keep_all = true, bool keep_dr = true, int sq_steps = 0) {
  dexsyn::Rng rng(seed);
  RefRunner run(gs);
  size_t nx = portElements(gs.prog.inputs());
  size_t nr = portElements(gs.prog.outputs());
  if (x.size() != nx) throw std::runtime_error("x size mismatch");
  std::vector<double> dx(nx);
  for (auto &v : dx) v = rng.uniform(-1, 1);
  bundle.arrays.emplace_back("x", x);
  bundle.arrays.emplace_back("dx", dx);
  std::vector<uint8_t> scalar_out;  // per output element: scalar-typed port?
  for (auto &p : gs.prog.outputs()) {
    uint8_t lanes, elem;
    classify(p.typeInfo(), lanes, elem);
    for (size_t i = 0; i < p.size() / sizeof(double); i++) scalar_out.push_back(lanes == 1);
  }
  std::vector<double> g_total, h_total;
  double loss_total = 0;
  std::vector<std::vector<double>> all_params;
  for (int t = 0; t < tiles; t++) {
    std::vector<double> params(n_param_lane * 8);
    for (size_t i = 0; i < params.size(); i++) {
      size_t port = i / 8;
      if (param_kind == "dexsyn") {
        // object start offsets U(-0.01,0.01), yaw U(0,2pi) (dexenv_grasp.h:151-164); dropout masks
        if (port == 0 || port == 1) params[i] = rng.uniform(-0.01, 0.01);
        else if (port == 2) params[i] = rng.uniform(0, 2 * M_PI);
        else params[i] = (rng.uniform() < 0.3) ? 0.0 : 1.0 / 0.7;
      } else {
        params[i] = rng.uniform(-0.1, 0.1);
      }
    }
    all_params.push_back(params);
    std::vector<double> r, v, g, dr, h;
    run.parameterize(params);
    run.forward(x, r);
    run.prepare();
    v = r;
    if (t > 0)
      for (size_t i = 0; i < v.size(); i++)
        if (scalar_out[i]) v[i] = 0;  // lane-independent residual rows count once
    run.reverse(v, g);
    run.tangent(dx, dr);
    run.gaussNewton(dx, h);
    for (size_t i = 0; i < r.size(); i++)
      if (!(t > 0 && scalar_out[i])) loss_total += r[i] * r[i];
    if (t == 0) {
      g_total = g;
      h_total = h;
    } else {
      for (size_t i = 0; i < g.size(); i++) g_total[i] += g[i];
      // hprop of a later tile re-propagates the scalar residual rows as well; remove them by
      // running bprop on the masked tangent instead
      std::vector<double> drm = dr, hm;
      for (size_t i = 0; i < drm.size(); i++)
        if (scalar_out[i]) drm[i] = 0;
      run.reverse(drm, hm);
      for (size_t i = 0; i < hm.size(); i++) h_total[i] += hm[i];
    }
    std::string s = "tile" + std::to_string(t) + "/";
    bundle.arrays.emplace_back(s + "params", params);
    bundle.arrays.emplace_back(s + "r", r);
    bundle.arrays.emplace_back(s + "g", g);
    if (keep_dr) bundle.arrays.emplace_back(s + "dr", dr);
  }
  bundle.arrays.emplace_back("g_total", g_total);
  bundle.arrays.emplace_back("h_total", h_total);
  bundle.arrays.emplace_back("loss_total", std::vector<double>{loss_total});
  (void)nr;

  // GradientDescentSolver::_step (tractor/include/tractor/solvers/gd.h:43-129) driven a fixed number
  // of times over all tiles: v <- mu v - lr g (mu = 0, dexopt.cpp:199-202), x <- accumulate(x, v)
  // through the reference's own accu program (solvers/base.h:158-165).
  if (gd_steps > 0) {
    std::vector<double> xk = x, losses, iterates;
    for (int step = 0; step < gd_steps; step++) {
      std::vector<double> g_sum(nx, 0.0);
      double loss = 0;
      for (int t = 0; t < tiles; t++) {
        std::vector<double> r, g;
        run.parameterize(all_params[t]);
        run.forward(xk, r);
        run.prepare();
        for (size_t i = 0; i < r.size(); i++) {
          if (t > 0 && scalar_out[i]) r[i] = 0;
          loss += r[i] * r[i];
        }
        run.reverse(r, g);
        for (size_t i = 0; i < nx; i++) g_sum[i] += g[i];
      }
      std::vector<double> ab(2 * nx), xn;
      for (size_t i = 0; i < nx; i++) {
        ab[i] = xk[i];
        ab[nx + i] = -learning_rate * g_sum[i];
      }
      run.x_accu->run(ab, run.memory, xn);
      xk = xn;
      losses.push_back(loss);
      if (gd_keep_all || step + 1 == gd_steps) iterates.insert(iterates.end(), xk.begin(), xk.end());
    }
    bundle.arrays.emplace_back("gd/loss", losses);
    bundle.arrays.emplace_back("gd/x", iterates);
    bundle.arrays.emplace_back("gd/learning_rate", std::vector<double>{learning_rate});
  }

  // LeastSquaresSolver::_step (tractor/include/tractor/solvers/sq.h:57-123) with dexopt's settings
  // (dexopt/src/dexopt.cpp:153-160: regularization 0.1, at most 100 linear iterations, step scaling 0.5, tolerance
  // 1e-9) driven a fixed number of times over all tiles.  The engine work -- x_prog, x_prep, x_bprop and the
  // Gauss-Newton product x_fprop ; x_bprop (= x_hprop, gradients.cpp:344-362; solvers/base.h:76-85 adds lambda v) --
  // is the reference's; the linear solver is a restatement of Eigen::ConjugateGradient<MatrixReplacement,
  // Lower|Upper, IdentityPreconditioner> (Eigen is not in this image: zero initial guess, stop when
  // |residual|^2 < tol^2 |rhs|^2, iteration cap) -- parity of the linear solve is pinned to this restatement only.
  if (sq_steps > 0) {
    const double lambda = 0.1, step_scaling = 0.5, tol = 1e-9;
    const int max_linear = 100;
    std::vector<std::unique_ptr<RefRunner>> runs;
    for (int t = 0; t < tiles; t++) {
      runs.emplace_back(new RefRunner(gs));
      runs.back()->parameterize(all_params[t]);
    }
    auto dot = [&](const std::vector<double> &a, const std::vector<double> &b) {
      double sum = 0;
      for (size_t i = 0; i < a.size(); i++) sum += a[i] * b[i];
      return sum;
    };
    // A v = sum over tiles of J_t^T J_t v (lane-independent rows once) + lambda v, at the current point
    auto product = [&](const std::vector<double> &v, std::vector<double> &out) {
      out.assign(nx, 0.0);
      for (int t = 0; t < tiles; t++) {
        std::vector<double> dr, h;
        runs[t]->tangent(v, dr);
        if (t > 0)
          for (size_t i = 0; i < dr.size(); i++)
            if (scalar_out[i]) dr[i] = 0;
        runs[t]->reverse(dr, h);
        for (size_t i = 0; i < nx; i++) out[i] += h[i];
      }
      for (size_t i = 0; i < nx; i++) out[i] += lambda * v[i];
    };
    std::vector<double> xk = x, losses, iterates, cg_its;
    for (int step = 0; step < sq_steps; step++) {
      std::vector<double> rhs(nx, 0.0);
      double loss = 0;
      for (int t = 0; t < tiles; t++) {
        std::vector<double> r, g;
        runs[t]->forward(xk, r);
        runs[t]->prepare();
        for (size_t i = 0; i < r.size(); i++) {
          if (t > 0 && scalar_out[i]) r[i] = 0;
          loss += r[i] * r[i];
        }
        runs[t]->reverse(r, g);
        for (size_t i = 0; i < nx; i++) rhs[i] += g[i];
      }
      // conjugate gradients
      std::vector<double> sol(nx, 0.0), residual = rhs, p, tmp;
      int it = 0;
      const double rhs_norm2 = dot(rhs, rhs);
      if (rhs_norm2 > 0) {
        const double threshold = std::max(tol * tol * rhs_norm2, std::numeric_limits<double>::min());
        double residual_norm2 = dot(residual, residual);
        if (residual_norm2 >= threshold) {
          p = residual;
          double abs_new = dot(residual, p);
          while (it < max_linear) {
            product(p, tmp);
            const double alpha = abs_new / dot(p, tmp);
            for (size_t i = 0; i < nx; i++) sol[i] += alpha * p[i];
            for (size_t i = 0; i < nx; i++) residual[i] -= alpha * tmp[i];
            residual_norm2 = dot(residual, residual);
            if (residual_norm2 < threshold) break;
            const double abs_old = abs_new;
            abs_new = dot(residual, residual);
            const double beta = abs_new / abs_old;
            for (size_t i = 0; i < nx; i++) p[i] = residual[i] + beta * p[i];
            it++;
          }
        }
      }
      std::vector<double> ab(2 * nx), xn;
      for (size_t i = 0; i < nx; i++) {
        ab[i] = xk[i];
        ab[nx + i] = -sol[i] * step_scaling;
      }
      runs[0]->x_accu->run(ab, runs[0]->memory, xn);
      xk = xn;
      losses.push_back(loss);
      cg_its.push_back((double)it);
      iterates.insert(iterates.end(), xk.begin(), xk.end());
    }
    bundle.arrays.emplace_back("sq/loss", losses);
    bundle.arrays.emplace_back("sq/x", iterates);
    bundle.arrays.emplace_back("sq/cg_iterations", cg_its);
    bundle.arrays.emplace_back("sq/settings", std::vector<double>{lambda, (double)max_linear, step_scaling, tol});
  }
}

static int cmdRecord(const std::string &path, const std::map<std::string, std::string> &kv) {
  std::string problem = getS(kv, "problem", "dexsyn");
  int tiles = (int)getI(kv, "tiles", 2);
  GradientSet gs;
  std::vector<double> x;
  size_t n_param_lane = 0;
  if (problem == "fkchain") {
    size_t nx;
    recordFkChain(gs, (int)getI(kv, "joints", 24), (uint64_t)getI(kv, "seed", 20211227), nx, n_param_lane);
    x.resize(nx);
    for (size_t i = 0; i < nx; i++) x[i] = 0.1 * (double)(i % 5) - 0.2 + 0.01 * (double)i;
  } else if (problem == "dexsyn") {
    recordDexSyn(gs, kv, x, n_param_lane);
  } else {
    throw std::runtime_error("unknown problem " + problem);
  }
  trxfile::Bundle bundle;
  std::string meta;
  for (auto &p : kv) meta += p.first + "=" + p.second + "\n";
  meta += "n_x=" + std::to_string(x.size()) + "\n";
  meta += "n_param_ports=" + std::to_string(n_param_lane) + "\n";
  meta += "n_outputs=" + std::to_string(gs.prog.outputs().size()) + "\n";
  {
    size_t n = 0;
    for (auto &i : gs.prog.instructions()) (void)i, n++;
    meta += "n_prog_instructions=" + std::to_string(n) + "\n";
  }
  bundle.texts.emplace_back("meta", meta);
  // programs=0: values only (the product records the same problem itself, trx_dexsyn_build);
  // gd_keep=last: store only the last gradient-descent iterate (all losses are kept);
  // keep_dr=0: drop the per-tile tangent residuals (h_total = J^T J dx stays)
  if (getI(kv, "programs", 1)) gs.addTo(bundle);
  if (tiles > 0)
    goldenRuns(gs, bundle, x, n_param_lane, tiles, (uint64_t)getI(kv, "seed", 20211227) + 77, problem,
               (int)getI(kv, "gd_steps", 0), getD(kv, "lr", 0.01), getS(kv, "gd_keep", "all") == "all",
               getI(kv, "keep_dr", 1) != 0, (int)getI(kv, "sq_steps", 0));
  else bundle.arrays.emplace_back("x", x);
  trxfile::writeBundle(bundle, path);
  std::cerr << meta;
  return 0;
}

// ------------------------------------------------------------------ CPU baseline

static int cmdBench(const std::map<std::string, std::string> &kv) {
  GradientSet gs;
  std::vector<double> x;
  size_t n_param_lane = 0;
  recordDexSyn(gs, kv, x, n_param_lane);
  int frames = (int)getI(kv, "frames", 4);
  int threads = (int)getI(kv, "threads", 1);
  double seconds = getD(kv, "seconds", 3.0);
  bool loop = getS(kv, "engine", "simple") == "loop";
  // Every thread first builds its own runner (own Memory + 6 compiled executables); the clock starts only when
  // all of them are ready, so the start-up cost is not charged to the rate.  `repeats` windows of `seconds`
  // each are measured back to back and the median window is reported.
  int repeats = (int)getI(kv, "repeats", 1);
  std::atomic<long> evals(0);
  std::atomic<bool> stop(false);
  std::atomic<int> ready(0);
  std::atomic<bool> go(false);
  std::vector<std::thread> pool;
  std::vector<double> checks(threads, 0.0);
  for (int ti = 0; ti < threads; ti++) {
    pool.emplace_back([&, ti]() {
      std::unique_ptr<RefRunner> runp;
      {
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        MuteCout mute;  // LoopEngine prints its schedule while compiling
        runp.reset(new RefRunner(gs, loop));  // own Memory + executables per thread
      }
      RefRunner &run = *runp;
      dexsyn::Rng rng(1000 + ti);
      std::vector<double> params(n_param_lane * 8), r, g;
      auto evaluate = [&]() {
        for (size_t i = 0; i < params.size(); i++) {
          size_t port = i / 8;
          params[i] = port < 2 ? rng.uniform(-0.01, 0.01) : (port == 2 ? rng.uniform(0, 6.28) : 1.0);
        }
        run.parameterize(params);
        run.forward(x, r);   // x_prog: input, execute, output
        run.prepare();       // x_prep
        run.reverse(r, g);   // x_bprop: input, execute, output   (gd.h:50-63)
        checks[ti] += g[0];
      };
      evaluate();  // first touch of the memory image is not timed either
      ready.fetch_add(1);
      while (!go.load()) std::this_thread::yield();
      while (!stop.load()) {
        evaluate();
        evals.fetch_add(1);
      }
    });
  }
  while (ready.load() < threads) std::this_thread::sleep_for(std::chrono::milliseconds(5));
  go.store(true);
  std::vector<double> rates;
  const double units = 8.0 * frames * getI(kv, "outer", 1);
  auto t_prev = std::chrono::steady_clock::now();
  long n_prev = evals.load();
  double dt_total = 0;
  for (int rep = 0; rep < repeats; rep++) {
    std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
    auto t_now = std::chrono::steady_clock::now();
    long n_now = evals.load();
    double dt = std::chrono::duration<double>(t_now - t_prev).count();
    rates.push_back((double)(n_now - n_prev) * units / dt);
    dt_total += dt;
    t_prev = t_now;
    n_prev = n_now;
  }
  stop.store(true);
  for (auto &t : pool) t.join();
  long n = n_prev;
  std::vector<double> sorted = rates;
  std::sort(sorted.begin(), sorted.end());
  double rate = sorted[sorted.size() / 2];
  std::printf("{\"evals\": %ld, \"seconds\": %.4f, \"threads\": %d, \"frames\": %d, \"lanes\": 8, "
              "\"step_rollouts_per_s\": %.3f, \"engine\": \"%s\", \"windows\": [",
              n, dt_total, threads, frames, rate, loop ? "loop" : "simple");
  for (size_t i = 0; i < rates.size(); i++) std::printf("%s%.1f", i ? ", " : "", rates[i]);
  std::printf("]}\n");
  return 0;
}

int main(int argc, char **argv) {
  try {
    if (argc < 2) throw std::runtime_error("usage: ref_tool registry|optests|record|bench ...");
    std::string cmd = argv[1];
    if (cmd == "registry") return cmdRegistry(argv[2]);
    if (cmd == "constraint_tests") return cmdConstraintTests(argv[2], parseArgs(argc, argv, 3));
    if (cmd == "constrained") return cmdConstrained(argv[2], parseArgs(argc, argv, 3));
    if (cmd == "optests") return cmdOpTests(argv[2], parseArgs(argc, argv, 3));
    if (cmd == "record") return cmdRecord(argv[2], parseArgs(argc, argv, 3));
    if (cmd == "bench") return cmdBench(parseArgs(argc, argv, 2));
    throw std::runtime_error("unknown command " + cmd);
  } catch (const std::exception &e) {
    std::cerr << "ref_tool: " << e.what() << std::endl;
    return 1;
  }
}
