// dexsyn_options.h -- "key=value key=value ..." option strings of the synthetic problems, shared by the
// flat-tape assembly (trx_builder.cpp), the re-rolled assembly (trx_rollout.cu) and the test emulation.
#pragma once
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "dexsyn_model.h"

namespace dexsyn {

struct ProblemOptions {
  ModelOptions model;
  int frames = 4;
  int outer = 1;
  double wstd = 0.1;
  std::map<std::string, std::string> kv;  // everything given, for the bundle's "meta" text
};

inline ProblemOptions parseProblemOptions(const char *options) {
  ProblemOptions po;
  auto &kv = po.kv;
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
  auto getS = [&](const char *k, const char *d) { return kv.count(k) ? kv[k] : std::string(d); };
  auto getI = [&](const char *k, long d) { return kv.count(k) ? std::atol(kv[k].c_str()) : d; };
  auto getD = [&](const char *k, double d) { return kv.count(k) ? std::atof(kv[k].c_str()) : d; };
  po.model.kind = getS("kind", "handarm");
  po.model.chain_joints = (int)getI("joints", 30);
  po.model.chain_contacts = (int)getI("contacts", 16);
  po.model.extra_fixed = (int)getI("extra_fixed", 12);
  po.model.policy_hidden = (int)getI("hidden", 64);
  po.model.dropout = getI("dropout", 0) != 0;
  po.model.collision = getI("collision", 0) != 0;
  po.model.friction = getI("friction", 0) != 0;
  po.model.floor_pair = getI("floor_pair", 1) != 0;
  po.model.seed = (uint64_t)getI("seed", 20211227);
  po.frames = (int)getI("frames", 4);
  po.outer = (int)getI("outer", 1);
  po.wstd = getD("wstd", 0.1);
  if (po.frames < 1) throw std::runtime_error("frames must be positive");
  return po;
}

// initial policy parameters: N(0, wstd^2), the same draws for every recorder (seed + 1)
inline std::vector<double> initialParameters(const ProblemOptions &po, size_t n) {
  Rng rng(po.model.seed + 1);
  std::vector<double> x0(n);
  for (auto &v : x0) v = po.wstd * rng.normal();
  return x0;
}

}  // namespace dexsyn
