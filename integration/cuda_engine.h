// cuda_engine.h -- tractor::Engine implementation backed by the B200 tape engine.
//
// This is the reference-side binding a tractor maintainer adds to switch dexopt from
//     auto engine = std::make_shared<tractor::SimpleEngine>();          (dexopt/src/dexopt.cpp:78)
// to
//     auto engine = std::make_shared<tractor::CudaEngine>();
// It derives from the reference's own interface (tractor/include/tractor/core/engine.h:11-112) and
// forwards the five private virtuals of Executable to the C ABI in include/trx_b200.h:
//     _compile       -> trx_program_create   (Program -> neutral description: Operator::name() + argument
//                                              layouts from Operator::arguments(), ports, constants)
//     _execute       -> trx_execute
//     _input         -> trx_input
//     _output        -> trx_output
//     _parameterize  -> trx_parameterize
// Errors of the C ABI are rethrown as std::runtime_error, the reference's error channel
// (core/engine.h:20-24, src/core/engine.cpp:44-48).
//
// Needs the tractor headers, so it is compiled only where the reference is available
// (oracle/ref_build/ref_cuda_engine_test.cpp builds and runs it against SimpleEngine).
#pragma once
#include <cxxabi.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <tractor/core/engine.h>
#include <tractor/core/operator.h>
#include <tractor/core/program.h>

#include "../include/trx_b200.h"

// Tapes with collision_axes / collision_project_2 (dexopt/src/ops.h:208-375) carry the addresses of host objects
// (CollisionShape<double>*, ShapeCollisionPair<double>*) as uint64 constants.  Where the reference's collision
// headers are available (they need Eigen; the queries themselves no longer need Bullet), define
// TRX_WITH_TRACTOR_COLLISION before including this file: _compile then registers every such object with the
// library's shape table (trx_collision_shape_create / _pair_create) and hands the device the handles instead.
// Without the macro such a tape is refused by trx_program_create ("not a handle").  This block cannot be compiled
// in an image without Eigen and is therefore exercised only through the C ABI (tests/test_collision.py).
#ifdef TRX_WITH_TRACTOR_COLLISION
#include <map>
#include <cstring>
#include <tractor/collision/query.h>
#include <tractor/collision/shape.h>
#endif

namespace tractor {

class CudaEngine : public Engine {
  int _device = 0;
  // rollouts per Memory.  8 = one Batch<double,8> image: every Buffer is byte-identical to the reference's, so
  // solvers, inputVector / outputVector and the tests work unchanged.  A multiple of 8 runs that many lane tiles per
  // launch; Buffers handed to input() / parameterize() / output() are then the wide packed layout of
  // include/trx_b200.h (port order kept, lane-typed ports [component][lanes], scalar-typed ports once; scalar
  // gradients of reverse programs summed over the tiles) -- the form the objective-level calls use.
  uint64_t _lanes = 8;

  static void check(trx_status s) {
    if (s != TRX_OK) throw std::runtime_error(std::string("CudaEngine: ") + trx_last_error());
  }

  static void classify(const TypeInfo &t, uint8_t &lanes, uint8_t &elem) {
    int status = 0;
    char *d = abi::__cxa_demangle(t.name(), nullptr, nullptr, &status);
    std::string n = (status == 0 && d) ? d : t.name();
    free(d);
    lanes = 1;
    elem = 0;
    auto pos = n.find("tractor::Batch<");
    if (pos != std::string::npos) lanes = (uint8_t)std::atoi(n.c_str() + n.find(',', pos) + 1);
    else if (n.find("unsigned long") != std::string::npos) elem = 1;
  }

  class MemoryImpl : public Memory {
   public:
    trx_memory *handle = nullptr;
    MemoryImpl(uint64_t lanes, int device) { check(trx_memory_create(lanes, device, &handle)); }
    ~MemoryImpl() override { trx_memory_destroy(handle); }
    MemoryImpl(const MemoryImpl &) = delete;
    MemoryImpl &operator=(const MemoryImpl &) = delete;
  };

  class ExecutableImpl : public Executable {
    int _device;
    uint64_t _lanes;
    trx_program *_program = nullptr;

    template <class Ports> static std::vector<trx_port_desc> ports(const Ports &src) {
      std::vector<trx_port_desc> out;
      for (auto &p : src) {
        trx_port_desc d{};
        d.address = p.address();
        d.offset = p.offset();
        d.size = (uint32_t)p.size();
        classify(p.typeInfo(), d.lanes, d.elem);
        out.push_back(d);
      }
      return out;
    }
    static trx_memory *mem(const std::shared_ptr<Memory> &m) { return static_cast<MemoryImpl *>(m.get())->handle; }

#ifdef TRX_WITH_TRACTOR_COLLISION
    // one table entry per host object, process-wide (the same shape is shared by many instructions and programs)
    static uint64_t shapeHandle(const CollisionShape<double> *shape) {
      static std::map<const void *, uint64_t> known;
      auto it = known.find(shape);
      if (it != known.end()) return it->second;
      trx_shape_desc d{};
      std::vector<double> points, planes;
      if (auto *cyl = dynamic_cast<const CollisionCylinderShape<double> *>(shape)) {
        d.kind = TRX_SHAPE_CYLINDER;
        d.radius = cyl->radius();
        d.length = cyl->length();
      } else if (auto *mesh = dynamic_cast<const ConvexPolyhedralCollisionShape<double> *>(shape)) {
        d.kind = TRX_SHAPE_POLYHEDRON;
        for (auto &p : mesh->points()) points.insert(points.end(), {p.x(), p.y(), p.z()});
        for (auto &pl : mesh->planes())
          planes.insert(planes.end(), {pl.normal().x(), pl.normal().y(), pl.normal().z(), pl.offset()});
        d.n_points = (uint32_t)mesh->points().size();
        d.n_planes = (uint32_t)mesh->planes().size();
        d.points = points.data();
        d.planes = planes.data();
      } else if (dynamic_cast<const CollisionSphereShape<double> *>(shape)) {
        // the radius has no accessor: support(+x) = centre + radius * x (shape.h:60-63)
        d.kind = TRX_SHAPE_SPHERE;
        auto c = shape->center();
        auto s = shape->support(Vector3<double>(1, 0, 0));
        d.center[0] = c.x(), d.center[1] = c.y(), d.center[2] = c.z();
        d.radius = s.x() - c.x();
      } else {
        throw std::runtime_error("CudaEngine: collision shape not yet supported");  // as dexlearn.h:212-216
      }
      uint64_t h = 0;
      check(trx_collision_shape_create(&d, &h));
      return known[shape] = h;
    }
    static uint64_t pairHandle(const ShapeCollisionPair<double> *pair) {
      static std::map<const void *, uint64_t> known;
      auto it = known.find(pair);
      if (it != known.end()) return it->second;
      uint64_t h = 0;
      check(trx_collision_pair_create(shapeHandle(pair->shapeA().get()), shapeHandle(pair->shapeB().get()), &h));
      return known[pair] = h;
    }
    // rewrites, in a copy of the constant blob, the uint64 constants that collision instructions read
    static std::vector<uint8_t> translateShapeConstants(const Program &program) {
      std::vector<uint8_t> blob(program.constData().begin(), program.constData().end());
      for (auto &inst : program.instructions()) {
        const std::string &name = inst.op()->name();
        const bool axes = name.rfind("collision_axes_", 0) == 0, proj = name.rfind("collision_project_2_", 0) == 0;
        if (!axes && !proj) continue;
        const uint64_t slot = (uint64_t)inst.args()[axes ? 2 : 1];
        for (auto &c : program.constants()) {
          if (c.address() != slot) continue;
          uint64_t bits = 0;
          std::memcpy(&bits, blob.data() + c.offset(), 8);
          const uint64_t h = axes ? pairHandle((const ShapeCollisionPair<double> *)bits)
                                  : shapeHandle((const CollisionShape<double> *)bits);
          std::memcpy(blob.data() + c.offset(), &h, 8);
        }
      }
      return blob;
    }
#endif

   protected:
    void _compile(const Program &program) override {
      std::vector<const Operator *> op_list;
      std::vector<std::vector<trx_arg_desc>> arg_store;
      std::vector<uint64_t> code;
      uint64_t n_inst = 0;
      for (auto &inst : program.instructions()) {
        const Operator *op = inst.op();
        size_t index = 0;
        while (index < op_list.size() && op_list[index] != op) index++;
        if (index == op_list.size()) {
          op_list.push_back(op);
          std::vector<trx_arg_desc> args;
          for (auto &a : op->arguments()) {
            trx_arg_desc d{};
            d.size = (uint32_t)a.size();
            d.is_input = a.isInput() ? 1 : 0;
            classify(a.typeInfo(), d.lanes, d.elem);
            args.push_back(d);
          }
          arg_store.push_back(std::move(args));
        }
        code.push_back(index);
        for (auto a : inst.args()) code.push_back((uint64_t)a);
        n_inst++;
      }
      std::vector<trx_op_desc> ops(op_list.size());
      for (size_t i = 0; i < ops.size(); i++) {
        ops[i].name = op_list[i]->name().c_str();
        ops[i].n_args = (uint32_t)arg_store[i].size();
        ops[i].args = arg_store[i].data();
      }
      auto in = ports(program.inputs()), par = ports(program.parameters()), out = ports(program.outputs()),
           con = ports(program.constants());
      trx_program_desc d{};
      d.memory_size = program.memorySize();
      d.n_ops = (uint32_t)ops.size();
      d.ops = ops.data();
      d.n_instructions = n_inst;
      d.n_code_words = code.size();
      d.code = code.data();
      d.n_inputs = (uint32_t)in.size();
      d.inputs = in.data();
      d.n_parameters = (uint32_t)par.size();
      d.parameters = par.data();
      d.n_outputs = (uint32_t)out.size();
      d.outputs = out.data();
      d.n_constants = (uint32_t)con.size();
      d.constants = con.data();
      d.n_const_data = program.constData().size();
      d.const_data = program.constData().data();
#ifdef TRX_WITH_TRACTOR_COLLISION
      const std::vector<uint8_t> translated = translateShapeConstants(program);
      d.const_data = translated.data();
#endif
      if (_program) trx_program_destroy(_program);
      _program = nullptr;
      check(trx_program_create(&d, _device, &_program));
    }
    void _execute(const std::shared_ptr<Memory> &memory) const override { check(trx_execute(_program, mem(memory), nullptr)); }
    void _input(const Buffer &input, const std::shared_ptr<Memory> &memory) const override {
      check(trx_input(_program, mem(memory), input.data(), input.size(), nullptr));
    }
    void _parameterize(const Buffer &data, const std::shared_ptr<Memory> &memory) const override {
      check(trx_parameterize(_program, mem(memory), data.data(), data.size(), nullptr));
    }
    void _output(const std::shared_ptr<const Memory> &memory, Buffer &output) const override {
      uint64_t bytes = outputBufferSize();  // 8 lanes: the reference's packed size
      if (_lanes != 8) check(trx_program_buffer_bytes(_program, TRX_BUF_OUTPUT, _lanes, &bytes));
      output.resize(bytes);
      auto *m = static_cast<const MemoryImpl *>(memory.get())->handle;
      check(trx_output(_program, m, output.data(), output.size(), nullptr));
    }

   public:
    ExecutableImpl(int device, uint64_t lanes) : _device(device), _lanes(lanes) {}
    ~ExecutableImpl() override { trx_program_destroy(_program); }
  };

 public:
  explicit CudaEngine(int device = 0, uint64_t lanes = 8) : _device(device), _lanes(lanes) {
    if (trx_device_count() <= 0) throw std::runtime_error("CudaEngine: no CUDA device (there is no CPU fallback)");
    if (lanes == 0 || lanes % 8) throw std::runtime_error("CudaEngine: lanes must be a positive multiple of 8");
  }
  std::shared_ptr<Memory> createMemory() override { return std::make_shared<MemoryImpl>(_lanes, _device); }
  std::shared_ptr<Executable> createExecutable() override { return std::make_shared<ExecutableImpl>(_device, _lanes); }
};

}  // namespace tractor
