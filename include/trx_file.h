// trx_file.h -- on-disk serialisation of tractor tape programs ("TRXB" bundles).
//
// A bundle is the neutral, engine-independent form of what the reference keeps
// in a tractor::Program (tractor/include/tractor/core/program.h:39-303): the
// instruction words (operator identity + byte addresses), the typed ports
// (inputs / parameters / outputs / constants), the goal list, the constant blob
// and the memory-image size.  Operators are identified by Operator::name()
// (tractor/include/tractor/core/operator.h:125, naming scheme :322-336,394-428)
// because that is the only identity that is stable across engines.
//
// The same file carries named f64 arrays (test vectors / golden outputs) so a
// fixture is one file.  Header-only, no dependencies; used by the reference-side
// dumper (oracle/ref_build), the CPU oracle (oracle/) and the product library.
//
// Layout (little endian):
//   "TRXB0001" u32 n_sections { u32 kind; str name; u64 bytes; payload }...
//   kind 1 = program, 2 = f64 array, 3 = utf-8 text
//   str = u16 length + bytes
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace trxfile {

enum ElemKind : uint8_t { ELEM_F64 = 0, ELEM_U64 = 1, ELEM_OTHER = 2 };

struct ArgDesc {
  uint32_t size = 0;     // bytes in the reference memory image
  uint8_t is_input = 0;  // Operator::Argument::isInput()
  uint8_t lanes = 1;     // 1 = scalar-typed, 8 = Batch<double,8>-typed, ...
  uint8_t elem = ELEM_F64;
};

struct OpDesc {
  std::string name;
  std::vector<ArgDesc> args;
};

struct Port {
  uint64_t address = 0;  // byte address in the memory image
  uint64_t offset = 0;   // byte offset in the packed port buffer
  uint32_t size = 0;
  uint8_t lanes = 1;
  uint8_t elem = ELEM_F64;
};

struct Goal {
  uint64_t port = 0;
  int32_t priority = 0;
};

struct Program {
  uint32_t lane_width = 1;  // widest lane type that appears (8 for *_8d tapes)
  uint64_t memory_size = 0;
  std::vector<OpDesc> ops;     // operator table (only operators that appear)
  uint64_t n_instructions = 0;
  std::vector<uint64_t> code;  // [op_index, arg0, arg1, ...] per instruction
  std::vector<Port> inputs, parameters, outputs, constants;
  std::vector<Goal> goals;
  std::vector<uint8_t> const_data;
};

struct Bundle {
  std::vector<std::pair<std::string, Program>> programs;
  std::vector<std::pair<std::string, std::vector<double>>> arrays;
  std::vector<std::pair<std::string, std::string>> texts;

  const Program *program(const std::string &n) const {
    for (auto &p : programs) if (p.first == n) return &p.second;
    return nullptr;
  }
  const std::vector<double> *array(const std::string &n) const {
    for (auto &p : arrays) if (p.first == n) return &p.second;
    return nullptr;
  }
};

// ---------------------------------------------------------------- writer

class Writer {
  std::vector<uint8_t> &_b;

 public:
  explicit Writer(std::vector<uint8_t> &b) : _b(b) {}
  void raw(const void *p, size_t n) {
    const uint8_t *c = (const uint8_t *)p;
    _b.insert(_b.end(), c, c + n);
  }
  template <class T> void pod(const T &v) { raw(&v, sizeof(T)); }
  void str(const std::string &s) {
    if (s.size() > 65535) throw std::runtime_error("trxfile: string too long");
    pod<uint16_t>((uint16_t)s.size());
    raw(s.data(), s.size());
  }
};

inline void encodePorts(Writer &w, const std::vector<Port> &ports) {
  w.pod<uint32_t>((uint32_t)ports.size());
  for (auto &p : ports) {
    w.pod<uint64_t>(p.address);
    w.pod<uint64_t>(p.offset);
    w.pod<uint32_t>(p.size);
    w.pod<uint8_t>(p.lanes);
    w.pod<uint8_t>(p.elem);
  }
}

inline void encodeProgram(const Program &p, std::vector<uint8_t> &out) {
  Writer w(out);
  w.pod<uint32_t>(p.lane_width);
  w.pod<uint64_t>(p.memory_size);
  w.pod<uint32_t>((uint32_t)p.ops.size());
  for (auto &op : p.ops) {
    w.str(op.name);
    w.pod<uint8_t>((uint8_t)op.args.size());
    for (auto &a : op.args) {
      w.pod<uint32_t>(a.size);
      w.pod<uint8_t>(a.is_input);
      w.pod<uint8_t>(a.lanes);
      w.pod<uint8_t>(a.elem);
    }
  }
  w.pod<uint64_t>(p.n_instructions);
  w.pod<uint64_t>((uint64_t)p.code.size());
  w.raw(p.code.data(), p.code.size() * sizeof(uint64_t));
  encodePorts(w, p.inputs);
  encodePorts(w, p.parameters);
  encodePorts(w, p.outputs);
  encodePorts(w, p.constants);
  w.pod<uint32_t>((uint32_t)p.goals.size());
  for (auto &g : p.goals) {
    w.pod<uint64_t>(g.port);
    w.pod<int32_t>(g.priority);
  }
  w.pod<uint64_t>((uint64_t)p.const_data.size());
  w.raw(p.const_data.data(), p.const_data.size());
}

inline void encodeBundle(const Bundle &b, std::vector<uint8_t> &out) {
  Writer w(out);
  w.raw("TRXB0001", 8);
  w.pod<uint32_t>((uint32_t)(b.programs.size() + b.arrays.size() + b.texts.size()));
  for (auto &t : b.texts) {
    w.pod<uint32_t>(3);
    w.str(t.first);
    w.pod<uint64_t>((uint64_t)t.second.size());
    w.raw(t.second.data(), t.second.size());
  }
  for (auto &p : b.programs) {
    std::vector<uint8_t> payload;
    encodeProgram(p.second, payload);
    w.pod<uint32_t>(1);
    w.str(p.first);
    w.pod<uint64_t>((uint64_t)payload.size());
    w.raw(payload.data(), payload.size());
  }
  for (auto &a : b.arrays) {
    w.pod<uint32_t>(2);
    w.str(a.first);
    w.pod<uint64_t>((uint64_t)(a.second.size() * sizeof(double)));
    w.raw(a.second.data(), a.second.size() * sizeof(double));
  }
}

inline void writeBundle(const Bundle &b, const std::string &path) {
  std::vector<uint8_t> out;
  encodeBundle(b, out);
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("trxfile: cannot open for writing: " + path);
  size_t n = std::fwrite(out.data(), 1, out.size(), f);
  std::fclose(f);
  if (n != out.size()) throw std::runtime_error("trxfile: short write: " + path);
}

// ---------------------------------------------------------------- reader

class Reader {
  const uint8_t *_p, *_e;

 public:
  Reader(const void *p, size_t n) : _p((const uint8_t *)p), _e((const uint8_t *)p + n) {}
  void raw(void *dst, size_t n) {
    if ((size_t)(_e - _p) < n) throw std::runtime_error("trxfile: truncated");
    std::memcpy(dst, _p, n);
    _p += n;
  }
  const uint8_t *skip(size_t n) {
    if ((size_t)(_e - _p) < n) throw std::runtime_error("trxfile: truncated");
    const uint8_t *r = _p;
    _p += n;
    return r;
  }
  template <class T> T pod() {
    T v;
    raw(&v, sizeof(T));
    return v;
  }
  std::string str() {
    uint16_t n = pod<uint16_t>();
    std::string s(n, '\0');
    raw(&s[0], n);
    return s;
  }
  bool done() const { return _p == _e; }
};

inline void decodePorts(Reader &r, std::vector<Port> &ports) {
  uint32_t n = r.pod<uint32_t>();
  ports.resize(n);
  for (auto &p : ports) {
    p.address = r.pod<uint64_t>();
    p.offset = r.pod<uint64_t>();
    p.size = r.pod<uint32_t>();
    p.lanes = r.pod<uint8_t>();
    p.elem = r.pod<uint8_t>();
  }
}

inline void decodeProgram(const void *data, size_t bytes, Program &p) {
  Reader r(data, bytes);
  p.lane_width = r.pod<uint32_t>();
  p.memory_size = r.pod<uint64_t>();
  p.ops.resize(r.pod<uint32_t>());
  for (auto &op : p.ops) {
    op.name = r.str();
    op.args.resize(r.pod<uint8_t>());
    for (auto &a : op.args) {
      a.size = r.pod<uint32_t>();
      a.is_input = r.pod<uint8_t>();
      a.lanes = r.pod<uint8_t>();
      a.elem = r.pod<uint8_t>();
    }
  }
  p.n_instructions = r.pod<uint64_t>();
  uint64_t nc = r.pod<uint64_t>();
  p.code.resize(nc);
  r.raw(p.code.data(), nc * sizeof(uint64_t));
  decodePorts(r, p.inputs);
  decodePorts(r, p.parameters);
  decodePorts(r, p.outputs);
  decodePorts(r, p.constants);
  p.goals.resize(r.pod<uint32_t>());
  for (auto &g : p.goals) {
    g.port = r.pod<uint64_t>();
    g.priority = r.pod<int32_t>();
  }
  uint64_t ncd = r.pod<uint64_t>();
  p.const_data.resize(ncd);
  r.raw(p.const_data.data(), ncd);
  if (!r.done()) throw std::runtime_error("trxfile: trailing bytes in program");
}

inline void decodeBundle(const void *data, size_t bytes, Bundle &b) {
  Reader r(data, bytes);
  char magic[8];
  r.raw(magic, 8);
  if (std::memcmp(magic, "TRXB0001", 8) != 0) throw std::runtime_error("trxfile: bad magic");
  uint32_t n = r.pod<uint32_t>();
  for (uint32_t i = 0; i < n; i++) {
    uint32_t kind = r.pod<uint32_t>();
    std::string name = r.str();
    uint64_t len = r.pod<uint64_t>();
    const uint8_t *payload = r.skip(len);
    if (kind == 1) {
      b.programs.emplace_back(name, Program());
      decodeProgram(payload, len, b.programs.back().second);
    } else if (kind == 2) {
      std::vector<double> v(len / sizeof(double));
      std::memcpy(v.data(), payload, v.size() * sizeof(double));
      b.arrays.emplace_back(name, std::move(v));
    } else if (kind == 3) {
      b.texts.emplace_back(name, std::string((const char *)payload, len));
    } else {
      throw std::runtime_error("trxfile: unknown section kind");
    }
  }
}

inline void readBundle(const std::string &path, Bundle &b) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("trxfile: cannot open: " + path);
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)n);
  size_t got = std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  if (got != buf.size()) throw std::runtime_error("trxfile: short read: " + path);
  decodeBundle(buf.data(), buf.size(), b);
}

}  // namespace trxfile
