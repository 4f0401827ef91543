// trx_builder.cpp -- C ABI of the product-side objective assembly ("dexopt's objective
// assembly" in BASELINE.json's north star): builds the synthetic dexopt rollout problem
// (csrc/assembly) onto a tape with the product's own recorder, differentiates it with the
// product's own buildGradients restatement, and returns everything as one TRXB bundle.
// Pure host code; the returned programs are executed by trx_engine.cu.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/trx_b200.h"
#include "assembly/dexsyn_assemble.h"
#include "builder/tape_builder.h"

namespace {
thread_local std::string g_builder_error;
}

// Registers the model's link shapes with the engine's collision shape table (what CollisionRobot does from the
// URDF meshes in the reference, src/collision/robot.cpp:60-140) and lists the link pairs to test: every end effector
// (finger tips, floor) against the object, and neighbouring finger tips against each other.
static void registerShapes(dexsyn::Model &model) {
  if (!model.wants_collision_pairs && !model.friction_cones) return;
  auto reg = [](dexsyn::Shape &s) {
    trx_shape_desc d;
    std::memset(&d, 0, sizeof(d));
    std::vector<double> planes;
    if (s.kind == 1) {
      d.kind = TRX_SHAPE_CYLINDER;
      d.radius = s.radius;
      d.length = s.length;
    } else {
      d.kind = TRX_SHAPE_POLYHEDRON;
      for (auto &pl : s.planes) planes.insert(planes.end(), {pl.n[0], pl.n[1], pl.n[2], pl.d});
      d.n_points = (uint32_t)(s.points.size() / 3);
      d.n_planes = (uint32_t)s.planes.size();
      d.points = s.points.data();
      d.planes = planes.data();
    }
    if (trx_collision_shape_create(&d, &s.handle) != TRX_OK) throw std::runtime_error(trx_last_error());
  };
  reg(model.object_shape);
  for (auto &s : model.ee_shapes) reg(s);
  if (!model.wants_collision_pairs) return;
  auto pair = [&](int ea, const dexsyn::Shape &sa, int link_b, const dexsyn::Shape &sb) {
    dexsyn::CollisionPair cp;
    cp.link_a = model.end_effectors[ea];
    cp.link_b = link_b;
    if (trx_collision_pair_create(sa.handle, sb.handle, &cp.handle) != TRX_OK) throw std::runtime_error(trx_last_error());
    model.collision_pairs.push_back(cp);
  };
  const int n_ee = (int)model.end_effectors.size();
  // (the floor is the last end effector)
  for (int e = 0; e < n_ee - (model.floor_pair ? 0 : 1); e++) pair(e, model.ee_shapes[e], model.object_joint, model.object_shape);
  for (int e = 0; e + 2 < n_ee; e++) pair(e, model.ee_shapes[e], model.end_effectors[e + 1], model.ee_shapes[e + 1]);
}

extern "C" {

const char *trx_builder_last_error(void) { return g_builder_error.c_str(); }

// options: "key=value" pairs separated by spaces; keys: kind frames outer hidden dropout seed
//          joints contacts extra_fixed wstd fuse_dense programs (comma separated subset of prog,prep,fprop,bprop,hprop,accu)
// On success *out points to a malloc'ed TRXB bundle (free with trx_builder_free) containing programs
// prog/prep/fprop/bprop/hprop/accu, array "x" (initial policy parameters) and text "meta".
int trx_dexsyn_build(const char *options, uint8_t **out, uint64_t *out_bytes) {
  try {
    std::map<std::string, std::string> kv;
    {
      std::string s = options ? options : "";
      size_t i = 0;
      while (i < s.size()) {
        while (i < s.size() && s[i] == ' ') i++;
        size_t j = s.find(' ', i);
        if (j == std::string::npos) j = s.size();
        std::string tok = s.substr(i, j - i);
        auto eq = tok.find('=');
        if (!tok.empty()) {
          if (eq == std::string::npos) throw std::runtime_error("expected key=value: " + tok);
          kv[tok.substr(0, eq)] = tok.substr(eq + 1);
        }
        i = j;
      }
    }
    auto getS = [&](const char *k, const char *d) { return kv.count(k) ? kv[k] : std::string(d); };
    auto getI = [&](const char *k, long d) { return kv.count(k) ? std::atol(kv[k].c_str()) : d; };
    auto getD = [&](const char *k, double d) { return kv.count(k) ? std::atof(kv[k].c_str()) : d; };

    dexsyn::ModelOptions mo;
    mo.kind = getS("kind", "handarm");
    mo.chain_joints = (int)getI("joints", 30);
    mo.chain_contacts = (int)getI("contacts", 16);
    mo.extra_fixed = (int)getI("extra_fixed", 12);
    mo.policy_hidden = (int)getI("hidden", 64);
    mo.dropout = getI("dropout", 0) != 0;
    mo.collision = getI("collision", 0) != 0;
    mo.friction = getI("friction", 0) != 0;
    mo.floor_pair = getI("floor_pair", 1) != 0;
    mo.seed = (uint64_t)getI("seed", 20211227);
    int frames = (int)getI("frames", 4);
    int outer = (int)getI("outer", 1);
    double wstd = getD("wstd", 0.1);
    dexsyn::Model model = dexsyn::makeModel(mo);
    registerShapes(model);
    size_t np = dexsyn::Assembler<trxb::TapeBuilder>::parameterCount(model);
    dexsyn::Rng rng(mo.seed + 1);
    std::vector<double> x0(np);
    for (auto &v : x0) v = wstd * rng.normal();

    trxb::TapeBuilder b;
    b.fuse_dense = getI("fuse_dense", 0) != 0;
    {
      dexsyn::Assembler<trxb::TapeBuilder> as(b, model);
      for (int o = 0; o < outer; o++) as.rollout(frames, x0);
    }
    uint64_t recorded = b.prog.n_instructions;
    b.finish();
    trxb::GradientPrograms g;
    std::string want = "," + getS("programs", "prog,prep,fprop,bprop,hprop,accu") + ",";
    bool want_tangent = want.find(",fprop,") != std::string::npos || want.find(",hprop,") != std::string::npos;
    trxb::buildGradients(b.prog, g, want_tangent);

    trxfile::Bundle bundle;
    std::string meta;
    for (auto &p : kv) meta += p.first + "=" + p.second + "\n";
    meta += "n_x=" + std::to_string(x0.size()) + "\n";
    meta += "n_param_ports=" + std::to_string(b.prog.parameters.size()) + "\n";
    meta += "n_outputs=" + std::to_string(b.prog.outputs.size()) + "\n";
    meta += "n_recorded_instructions=" + std::to_string(recorded) + "\n";
    meta += "n_prog_instructions=" + std::to_string(b.prog.n_instructions) + "\n";
    meta += "n_controlled=" + std::to_string(model.controlled.size()) + "\n";
    meta += "n_joints=" + std::to_string(model.joints.size()) + "\n";
    meta += "n_end_effectors=" + std::to_string(model.end_effectors.size()) + "\n";
    bundle.texts.emplace_back("meta", meta);
    auto put = [&](const char *name, trxfile::Program &p) {
      if (want.find(std::string(",") + name + ",") != std::string::npos) bundle.programs.emplace_back(name, std::move(p));
    };
    put("prog", b.prog);
    put("prep", g.prep);
    put("fprop", g.fprop);
    put("bprop", g.bprop);
    put("hprop", g.hprop);
    put("accu", g.accu);
    bundle.arrays.emplace_back("x", x0);
    std::vector<uint8_t> bytes;
    trxfile::encodeBundle(bundle, bytes);
    *out = (uint8_t *)std::malloc(bytes.size());
    if (!*out) throw std::runtime_error("out of memory");
    std::memcpy(*out, bytes.data(), bytes.size());
    *out_bytes = bytes.size();
    return TRX_OK;
  } catch (const std::exception &e) {
    g_builder_error = e.what();
    return TRX_ERR_INVALID;
  }
}

void trx_builder_free(uint8_t *p) { std::free(p); }

}  // extern "C"
