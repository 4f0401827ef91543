// vm_emu.cpp -- TEST-ONLY host emulation of the B200 tape VM.
//
// Compiles the *same* lowering (csrc/trx_lower.h) and the *same* operator bodies
// (csrc/vm_ops.def through csrc/vm_exec.cuh, built as plain C++) into a small
// shared object that walks the per-warp instruction streams sequentially.  It
// exists so that the CPU test suite (pytest -m "not gpu") can check the
// lowering (levels, bundles, port maps) and every operator body against the
// reference's golden vectors without a GPU.  It is NOT part of the product:
// nothing in robio_2021_dexopt_b200/ loads it, libtrx_b200.so has no CPU path.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/trx_b200.h"
#include "../../robio_2021_dexopt_b200/csrc/trx_lower.h"
#include "../../robio_2021_dexopt_b200/csrc/trx_fuse.h"
#include "../../robio_2021_dexopt_b200/csrc/collision/trx_shape_table.h"
#include "../../robio_2021_dexopt_b200/csrc/trx_direct.h"
#include "../../robio_2021_dexopt_b200/csrc/vm_exec.cuh"

namespace {
thread_local std::string g_err;
int fail(int code, const std::string &m) {
  g_err = m;
  return code;
}
}  // namespace

struct emu_program {
  trx::LoweredProgram low;
  int scalar_inputs, scalar_outputs;
  const trx::PortLayout &layout(int which) const {
    return which == TRX_BUF_INPUT ? low.inputs : which == TRX_BUF_PARAMETER ? low.parameters : low.outputs;
  }
};
struct emu_memory {
  uint64_t lanes, n_tiles, tile_stride = 0;  // doubles
  std::vector<double> mem;
  void ensure(uint64_t bytes) {
    uint64_t need = (bytes + 255) / 256 * 256 / 8;
    if (need <= tile_stride) return;
    std::vector<double> fresh(need * n_tiles, 0.0);
    for (uint64_t t = 0; t < n_tiles; t++)
      std::memcpy(fresh.data() + t * need, mem.data() + t * tile_stride, tile_stride * 8);
    mem.swap(fresh);
    tile_stride = need;
  }
};

extern "C" {
const char *emu_last_error(void) { return g_err.c_str(); }

int emu_program_create_from_blob(const void *blob, size_t bytes, int warps, emu_program **out) {
  *out = nullptr;
  try {
    trxfile::Program src;
    trxfile::decodeProgram(blob, bytes, src);
    std::unique_ptr<emu_program> p(new emu_program());
    trx::LowerOptions opt;
    opt.n_warps = (uint32_t)warps;
    if (const char *w = getenv("TRX_CACHE_SLOTS")) opt.ring_slots = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_POOL_SLOTS")) opt.pool_slots = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_CHAIN")) opt.chain_cap = (uint32_t)atoi(w);
    if (const char *w = getenv("TRX_FOLD")) opt.fold_accumulate = atoi(w) != 0;
    trx::lowerProgram(src, opt, p->low);
    bool reverse_only = p->low.has_reverse && !p->low.has_forward;
    p->scalar_inputs = reverse_only ? TRX_SCALAR_REDUCE : TRX_SCALAR_SHARED;
    p->scalar_outputs = p->low.has_reverse ? TRX_SCALAR_REDUCE : TRX_SCALAR_SHARED;
    *out = p.release();
    return TRX_OK;
  } catch (const std::exception &e) {
    std::string m = e.what();
    return fail(m.find("unsupported operator") != std::string::npos ? TRX_ERR_UNSUPPORTED : TRX_ERR_INVALID, m);
  }
}
// the objective's fused lowering (csrc/trx_fuse.h): forward+prepare and adjoint
int emu_fused_create(const void *prog, size_t prog_n, const void *prep, size_t prep_n, const void *bprop, size_t bprop_n,
                     int warps, int ring_slots, int sync_penalty, emu_program **fwd, emu_program **adj) {
  *fwd = *adj = nullptr;
  try {
    trxfile::Program p_prog, p_prep, p_bprop;
    trxfile::decodeProgram(prog, prog_n, p_prog);
    trxfile::decodeProgram(prep, prep_n, p_prep);
    trxfile::decodeProgram(bprop, bprop_n, p_bprop);
    std::unique_ptr<emu_program> f(new emu_program()), a(new emu_program());
    trx::FuseOptions opt;
    opt.n_warps = (uint32_t)warps;
    opt.ring_slots = (uint32_t)ring_slots;
    opt.sync_penalty = (uint32_t)sync_penalty;
    trx::lowerObjective(p_prog, p_prep, p_bprop, opt, f->low, a->low);
    f->scalar_inputs = f->scalar_outputs = TRX_SCALAR_SHARED;
    a->scalar_inputs = a->scalar_outputs = TRX_SCALAR_REDUCE;
    *fwd = f.release();
    *adj = a.release();
    return TRX_OK;
  } catch (const std::exception &e) {
    return fail(TRX_ERR_INVALID, e.what());
  }
}
// objective lowering step of the product (trx_direct.h) on serialised tapes: the adjoint reads forward
// operands in place.  Returns malloc'ed blobs of the rewritten prep and bprop programs; stats3 =
// {prepare instructions dropped, arguments redirected, reverse_mul rewritten}.
int emu_direct_prepare(const void *prep, size_t prep_n, const void *bprop, size_t bprop_n, void **out_prep,
                       size_t *out_prep_n, void **out_bprop, size_t *out_bprop_n, uint64_t *stats3) {
  try {
    trxfile::Program p_prep, p_bprop;
    trxfile::decodeProgram(prep, prep_n, p_prep);
    trxfile::decodeProgram(bprop, bprop_n, p_bprop);
    trx::DirectStats st;
    trx::directPrepare(p_prep, p_bprop, st);
    std::vector<uint8_t> a, b;
    trxfile::encodeProgram(p_prep, a);
    trxfile::encodeProgram(p_bprop, b);
    *out_prep = std::malloc(a.size());
    *out_bprop = std::malloc(b.size());
    std::memcpy(*out_prep, a.data(), a.size());
    std::memcpy(*out_bprop, b.data(), b.size());
    *out_prep_n = a.size();
    *out_bprop_n = b.size();
    stats3[0] = st.n_prepare_dropped;
    stats3[1] = st.n_arguments_redirected;
    stats3[2] = st.n_mul_rewritten;
  } catch (const std::exception &e) {
    return fail(TRX_ERR_INVALID, e.what());
  }
  return TRX_OK;
}
void emu_free(void *p) { std::free(p); }

void emu_program_destroy(emu_program *p) { delete p; }
int emu_program_buffer_bytes(const emu_program *p, int which, uint64_t lanes, uint64_t *bytes) {
  *bytes = p->layout(which).wideElems(lanes) * 8;
  return TRX_OK;
}
int emu_program_stat(const emu_program *p, const char *key, uint64_t *value) {
  std::string k = key;
  if (k == "n_instructions") *value = p->low.n_instructions;
  else if (k == "n_levels") *value = p->low.n_levels;
  else if (k == "n_bundles") *value = p->low.n_bundles;
  else if (k == "stream_words") *value = p->low.stream.size();
  else if (k == "memory_size") *value = p->low.memory_size;
  else if (k == "max_level_width") *value = p->low.max_level_width;
  else if (k == "arg_bytes_per_tile") *value = p->low.arg_bytes_per_tile;
  else if (k == "n_ring_reads") *value = p->low.n_ring_reads;
  else if (k == "n_image_reads") *value = p->low.n_image_reads;
  else if (k == "n_ring_only_writes") *value = p->low.n_ring_only_writes;
  else if (k == "n_dual_writes") *value = p->low.n_dual_writes;
  else if (k == "n_image_writes") *value = p->low.n_image_writes;
  else if (k == "n_folded") *value = p->low.n_folded;
  else if (k == "n_chained") *value = p->low.n_chained;
  else return fail(TRX_ERR_INVALID, "unknown statistic");
  return TRX_OK;
}
// walks every warp's stream: each must start on a page boundary, no record may straddle a page, every
// stream must hold the same number of BARRIER words and end with END
int emu_program_check_streams(const emu_program *p) {
  const auto &low = p->low;
  const uint32_t PG = low.page_words;
  uint64_t barriers0 = 0;
  for (uint32_t w = 0; w < low.n_warps; w++) {
    uint64_t begin = low.stream_begin[w], end = low.stream_begin[w + 1], at = begin, barriers = 0;
    if (PG && (begin % PG || end % PG)) return fail(TRX_ERR_INVALID, "stream not page aligned");
    bool ended = false;
    while (at < end) {
      uint32_t hdr = low.stream[at], op = hdr & 0x7ffu;
      uint64_t n = 1;
      if (op == trx::OP_END) {
        ended = true;
        break;
      } else if (op == trx::OP_BARRIER) {
        barriers++;
      } else if (op == trx::OP_PAGE) {
        if (!PG) return fail(TRX_ERR_INVALID, "page word in an unpaged stream");
        n = PG - ((at - begin) % PG);
      } else if (op == trx::OP_PREFETCH) {
        n = 33;
      } else {
        uint32_t nargs = (hdr >> 17) & 15u;
        bool wide = (hdr >> 11) & 1u, dual = (hdr >> 22) & 1u;
        n = ((hdr >> 21) & 1u) ? 1 + nargs + 2 : 1 + (uint64_t)(dual ? 2 * nargs : nargs) * (wide ? 32 : 4);
      }
      if (PG && op != trx::OP_PAGE && (at - begin) / PG != (at - begin + n - 1) / PG)
        return fail(TRX_ERR_INVALID, "record straddles a stream page");
      at += n;
    }
    if (!ended) return fail(TRX_ERR_INVALID, "stream without END");
    if (w == 0) barriers0 = barriers;
    else if (barriers != barriers0) return fail(TRX_ERR_INVALID, "streams disagree on the number of levels");
  }
  if (barriers0 != low.n_levels) return fail(TRX_ERR_INVALID, "barrier count differs from the level count");
  return TRX_OK;
}

int emu_memory_create(uint64_t lanes, emu_memory **out) {
  if (lanes == 0 || lanes % 8) return fail(TRX_ERR_INVALID, "lanes must be a positive multiple of 8");
  auto *m = new emu_memory();
  m->lanes = lanes;
  m->n_tiles = lanes / 8;
  *out = m;
  return TRX_OK;
}
void emu_memory_destroy(emu_memory *m) { delete m; }
int emu_memory_peek(emu_memory *m, uint64_t tile, uint64_t address, void *dst, uint64_t bytes) {
  if (tile >= m->n_tiles || address + bytes > m->tile_stride * 8) return fail(TRX_ERR_INVALID, "peek out of range");
  std::memcpy(dst, (const uint8_t *)(m->mem.data() + tile * m->tile_stride) + address, bytes);
  return TRX_OK;
}

static int scatter(const emu_program *p, emu_memory *m, int which, const double *wide, uint64_t bytes) {
  const trx::PortLayout &L = p->layout(which);
  if (bytes != L.wideElems(m->lanes) * 8) return fail(TRX_ERR_SIZE, "packed buffer size mismatch");
  m->ensure(p->low.memory_size);
  bool once = which == TRX_BUF_INPUT && p->scalar_inputs == TRX_SCALAR_REDUCE;
  for (uint64_t t = 0; t < m->n_tiles; t++)
    for (auto &e : L.elems) {
      double v;
      if (e.lane == 0xff) {
        v = wide[(uint64_t)e.fixed + (uint64_t)e.per_lane * m->lanes];
        if (once && t != 0) v = 0.0;
      } else {
        v = wide[(uint64_t)e.fixed + (uint64_t)e.per_lane * m->lanes + t * 8 + e.lane];
      }
      m->mem[t * m->tile_stride + e.idx] = v;
    }
  return TRX_OK;
}
int emu_input(const emu_program *p, emu_memory *m, const void *packed, uint64_t bytes) {
  return scatter(p, m, TRX_BUF_INPUT, (const double *)packed, bytes);
}
int emu_parameterize(const emu_program *p, emu_memory *m, const void *packed, uint64_t bytes) {
  return scatter(p, m, TRX_BUF_PARAMETER, (const double *)packed, bytes);
}
int emu_output(const emu_program *p, emu_memory *m, void *packed, uint64_t bytes) {
  const trx::PortLayout &L = p->low.outputs;
  if (bytes != L.wideElems(m->lanes) * 8) return fail(TRX_ERR_SIZE, "packed buffer size mismatch");
  double *wide = (double *)packed;
  bool sum = p->scalar_outputs == TRX_SCALAR_REDUCE;
  for (uint64_t t = 0; t < m->n_tiles; t++)
    for (auto &e : L.elems) {
      if (e.lane == 0xff) {
        if (t != 0) continue;
        double v = m->mem[e.idx];
        if (sum)
          for (uint64_t u = 1; u < m->n_tiles; u++) v += m->mem[u * m->tile_stride + e.idx];
        wide[(uint64_t)e.fixed + (uint64_t)e.per_lane * m->lanes] = v;
      } else {
        wide[(uint64_t)e.fixed + (uint64_t)e.per_lane * m->lanes + t * 8 + e.lane] = m->mem[t * m->tile_stride + e.idx];
      }
    }
  return TRX_OK;
}

// Walks the W streams level by level, the warps of a level one after the other and the
// 32 threads of a warp in turn -- any valid schedule must give the same answer.
static unsigned long long g_emu_seed = 0x9E3779B97F4A7C15ull, g_emu_epoch = 0;
void emu_set_random_seed(unsigned long long seed) {
  g_emu_seed = seed;
  g_emu_epoch = 0;
}
// known-answer access to the generator of the in-tape random operators (csrc/vm_exec.cuh)
void emu_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  const uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
  uint32_t o[4];
  trxvm::philox4x32_10(c, k, o);
  for (int i = 0; i < 4; i++) out[i] = o[i];
}

// collision shapes: the emulator's own table (same record builder as the library), and the two queries on the host
static trxcol::ShapeTable g_emu_shapes;
int emu_collision_shape_create(const trx_shape_desc *d, uint64_t *handle) {
  bool unsupported = false;
  const std::string err = g_emu_shapes.addShape(*d, *handle, unsupported);
  trxvm::shapeHeapHost() = g_emu_shapes.heap.data();
  return err.empty() ? 0 : fail(unsupported ? -2 : -1, err);
}
int emu_collision_pair_create(uint64_t a, uint64_t b, uint64_t *handle) {
  const std::string err = g_emu_shapes.addPair(a, b, *handle);
  trxvm::shapeHeapHost() = g_emu_shapes.heap.data();
  return err.empty() ? 0 : fail(-1, err);
}
void emu_collision_clear() {
  g_emu_shapes.clear();
  trxvm::shapeHeapHost() = g_emu_shapes.heap.data();
}
// a copy of the library's table (trx_collision_table), so that tapes built by the product recorder find their handles
void emu_collision_set_table(const double *heap, uint64_t n) {
  g_emu_shapes.clear();
  g_emu_shapes.heap.assign(heap, heap + n);
  for (uint64_t at = 1; at < n;) {  // walk the records to list the handles
    if ((int)heap[at] == trxcol::SHAPE_PAIR) {
      g_emu_shapes.pairs.push_back(at);
      at += trxcol::PAIR_SIZE;
    } else {
      g_emu_shapes.shapes.push_back(at);
      at += trxcol::REC_DATA + 3 * (uint64_t)heap[at + trxcol::REC_NPOINTS] + 4 * (uint64_t)heap[at + trxcol::REC_NPLANES];
    }
  }
  trxvm::shapeHeapHost() = g_emu_shapes.heap.data();
}
// out = n x 12: a, b, normal, distance, gjk iterations, epa iterations
int emu_collision_query_pairs(uint64_t pair, const double *poses, uint64_t n, double *out) {
  if (!g_emu_shapes.isPair(pair)) return fail(-1, "unknown handle");
  for (uint64_t i = 0; i < n; i++) {
    const double *q = poses + 14 * i;
    trxvm::PoseV a, b;
    a.t = trxvm::V3{q[0], q[1], q[2]}, a.r = trxvm::Q4{q[3], q[4], q[5], q[6]};
    b.t = trxvm::V3{q[7], q[8], q[9]}, b.r = trxvm::Q4{q[10], q[11], q[12], q[13]};
    trxcol::Result r;
    trxcol::collideShapes(g_emu_shapes.heap.data(), pair, a, b, r);
    double *o = out + 12 * i;
    o[0] = r.a.x, o[1] = r.a.y, o[2] = r.a.z, o[3] = r.b.x, o[4] = r.b.y, o[5] = r.b.z;
    o[6] = r.n.x, o[7] = r.n.y, o[8] = r.n.z, o[9] = r.d, o[10] = r.gjk_iterations, o[11] = r.epa_iterations;
  }
  return 0;
}
int emu_collision_project_points(uint64_t shape, const double *points, uint64_t n, double *out) {
  if (!g_emu_shapes.isShape(shape)) return fail(-1, "unknown handle");
  for (uint64_t i = 0; i < n; i++) {
    trxvm::V3 nrm = trxvm::V3{0.0, 0.0, 0.0};
    double d = 0.0;
    trxcol::projectPoint(g_emu_shapes.heap.data(), shape, trxvm::V3{points[3 * i], points[3 * i + 1], points[3 * i + 2]}, nrm, d);
    out[4 * i] = nrm.x, out[4 * i + 1] = nrm.y, out[4 * i + 2] = nrm.z, out[4 * i + 3] = d;
  }
  return 0;
}

int emu_execute(const emu_program *p, emu_memory *m) {
  m->ensure(p->low.memory_size);
  const auto &low = p->low;
  if (low.draws_random) {
    trxvm::RngState &st = trxvm::rngHostState();
    st.seed = g_emu_seed;
    st.epoch = ++g_emu_epoch;
    st.base = m->mem.data();
    st.tile_stride = m->tile_stride;
    st.lane_base = 0;
  }
  for (uint64_t t = 0; t < m->n_tiles; t++) {
    double *mem = m->mem.data() + t * m->tile_stride;
    for (size_t i = 0; i < low.const_idx.size(); i++) std::memcpy(mem + low.const_idx[i], &low.const_bits[i], 8);
    uint32_t W = low.n_warps;
    std::vector<double> rings((size_t)W * std::max<uint32_t>(low.ring_doubles, 1), 0.0);
    if (low.cta_ring)
      for (size_t i = 0; i < low.pool_values.size(); i++) rings[low.pool_offset_doubles + i] = low.pool_values[i];
    std::vector<const uint32_t *> pc(W);
    for (uint32_t w = 0; w < W; w++) pc[w] = low.stream.data() + low.stream_begin[w];
    bool done = false;
    while (!done) {
      for (uint32_t w = 0; w < W; w++) {
        while (true) {
          uint32_t hdr = *pc[w];
          uint32_t op = hdr & 0x7ffu;
          if (op == trx::OP_END) {
            done = true;
            break;
          }
          if (op == trx::OP_BARRIER) {
            pc[w]++;
            break;
          }
          if (op == trx::OP_PREFETCH) {
            pc[w] += 33;
            continue;
          }
          if (op == trx::OP_PAGE) {
            size_t at = (size_t)(pc[w] - (low.stream.data() + low.stream_begin[w]));
            pc[w] += low.page_words - (at & (low.page_words - 1));
            continue;
          }
          bool wide = (hdr >> 11) & 1u;
          int nsub = (int)((hdr >> 12) & 31u) + 1;
          int nargs = (int)((hdr >> 17) & 15u);
          if ((hdr >> 21) & 1u) {
            // block-cooperative record: this warp's 32 threads do their share
            for (int lane32 = 0; lane32 < 32; lane32++)
              trxvm::exec_coop(op, pc[w] + 1, mem, (int)(w * 32 + lane32), (int)(W * 32));
            pc[w] += 1 + nargs + 2;
            continue;
          }
          bool dual = (hdr >> 22) & 1u;
          const bool plain = !low.cta_ring && !low.global_pool && low.ring_doubles == 0;
          for (int lane32 = 0; lane32 < 32; lane32++) {
            auto run = [&](auto &c) {
              c.m = mem;
              c.ring = low.cta_ring ? rings.data() : rings.data() + (size_t)w * low.ring_doubles;
              c.nargs = nargs;
              c.dual = dual;
              c.preloaded = false;
              c.gpool = low.global_pool ? low.pool_values.data() : nullptr;
              c.rings = !low.global_pool;
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
              c.a = pc[w] + 1 + sub;
              if (sub < nsub) trxvm::exec(op, c);
            };
            if (plain) {  // the staged kernel's argument-word format (bit 31 of an output = accumulate)
              trxvm::CtxPlain c;
              run(c);
            } else {
              trxvm::Ctx c;
              run(c);
            }
          }
          pc[w] += 1 + (dual ? 2 * nargs : nargs) * (wide ? 32 : 4);
        }
      }
    }
  }
  return TRX_OK;
}
}
