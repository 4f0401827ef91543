// trx_ops.h -- host-side view of the device operator vocabulary (vm_ops.def).
//
// Maps tape operator names (Operator::name(), tractor/include/tractor/core/
// operator.h:125, scheme :322-336) to device opcodes and checks the argument
// layout the tape carries against the device operator's signature.  An unknown
// name is a hard error: there is no CPU fallback behind this engine.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace trx {

enum Opcode : uint16_t {
#define TRX_OP(ENUM, NAME, MODE, SIG, ...) OP_##ENUM,
#include "vm_ops.def"
#undef TRX_OP
  OP_COUNT,
  OP_PAGE = 0x7fc,      // continue at the next stream page (records never straddle a page)
  OP_PREFETCH = 0x7fd,  // 32 words: 128-byte line indices of the tile image to pull into L2
  OP_BARRIER = 0x7fe,
  OP_END = 0x7ff
};

// ARG_T lane-typed value, ARG_S BatchScalar, ARG_U uint64 handle; block kinds (components = 0 in the
// signature, the size comes from the program's operator table): ARG_TB contiguous lane-typed slots,
// ARG_SB contiguous scalar slots, ARG_SA scalar block that the operator accumulates into (+=)
enum ArgKind : uint8_t { ARG_T = 0, ARG_S = 1, ARG_U = 2, ARG_TB = 3, ARG_SB = 4, ARG_SA = 5 };

struct ArgSig {
  bool is_input;
  ArgKind kind;
  uint8_t comps;
};

struct OpSig {
  const char *base;  // e.g. "mul_fg_pose"
  char mode;         // C P F R
  const char *sig;   // "iT7 iT7 oT7"
  std::vector<ArgSig> args;
  std::string tape_name;  // "[mode_]base"
  unsigned cost = 0;      // components moved, used for load balancing
  bool coop = false;      // block-cooperative operator (has block arguments)
};

inline const std::vector<OpSig> &opTable() {
  static const std::vector<OpSig> table = [] {
    std::vector<OpSig> t;
#define TRX_OP(ENUM, NAME, MODE, SIG, ...) t.push_back(OpSig{NAME, #MODE[0], SIG, {}, "", 0});
#include "vm_ops.def"
#undef TRX_OP
    for (auto &op : t) {
      const char *p = op.sig;
      while (*p) {
        while (*p == ' ') p++;
        if (!*p) break;
        ArgSig a;
        a.is_input = (*p++ == 'i');
        char k = *p++;
        a.kind = (k == 'T') ? ARG_T : (k == 'S') ? ARG_S : (k == 'B') ? ARG_TB : (k == 'C') ? ARG_SB : (k == 'A') ? ARG_SA : ARG_U;
        int n = 0;
        while (*p >= '0' && *p <= '9') n = n * 10 + (*p++ - '0');
        a.comps = (uint8_t)n;
        op.args.push_back(a);
        op.cost += (unsigned)n;
        if (a.kind >= ARG_TB) op.coop = true;
      }
      // C P F R, and the constraint modes J I K D X Y Z (vm_ops.def, constraint operators)
      const char *prefix = op.mode == 'P' ? "prepare_" : op.mode == 'F' ? "forward_" : op.mode == 'R' ? "reverse_"
                           : op.mode == 'J' ? "project_" : op.mode == 'I' ? "barrier_init_" : op.mode == 'K' ? "barrier_step_"
                           : op.mode == 'D' ? "barrier_diagonal_" : op.mode == 'X' ? "penalty_init_"
                           : op.mode == 'Y' ? "penalty_step_" : op.mode == 'Z' ? "penalty_diagonal_" : "";
      op.tape_name = std::string(prefix) + op.base;
    }
    return t;
  }();
  return table;
}

inline bool isBlock(ArgKind k) { return k >= ARG_TB; }
// bytes per element of a block argument
inline uint32_t blockElemBytes(ArgKind k, uint32_t lane_width) { return k == ARG_TB ? 8 * lane_width : 8; }

// "prepare_mul_fg_pose_8d" -> (opcode, lane width); returns false if unknown
inline bool findOp(const std::string &tape_name, uint16_t &opcode, uint32_t &lane_width) {
  static const std::map<std::string, uint16_t> index = [] {
    std::map<std::string, uint16_t> m;
    auto &t = opTable();
    for (size_t i = 0; i < t.size(); i++) m[t[i].tape_name] = (uint16_t)i;
    return m;
  }();
  auto us = tape_name.rfind('_');
  if (us == std::string::npos) return false;
  std::string suffix = tape_name.substr(us + 1);
  if (suffix == "d") lane_width = 1;
  else if (suffix == "8d") lane_width = 8;
  else return false;
  auto it = index.find(tape_name.substr(0, us));
  if (it == index.end()) return false;
  opcode = it->second;
  return true;
}

}  // namespace trx
