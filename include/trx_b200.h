/* trx_b200.h -- C ABI of the B200 tape execution engine.
 *
 * Drop-in boundary: everything a tractor::Engine implementation needs in order
 * to execute tractor::Program objects on a B200 (sm_100a) instead of the CPU
 * interpreter.  Each entry point names the reference interface it stands behind
 * (paths relative to the reference repository):
 *
 *   reference (C++)                                         this ABI
 *   ------------------------------------------------------  --------------------------
 *   Engine::createMemory()            core/engine.h:109      trx_memory_create
 *   Engine::createExecutable()+compile core/engine.h:110-111,
 *     Executable::_compile            core/engine.h:25,
 *     SimpleEngine::_compile          src/engines/simple.cpp:7-20   trx_program_create[_from_blob]
 *   Executable::_execute              core/engine.h:26,
 *     SimpleEngine::_execute          src/engines/simple.cpp:22-35  trx_execute
 *   Executable::_input                core/engine.h:27-28,
 *     ExecutableBase::_input          src/engines/base.cpp:20-31    trx_input
 *   Executable::_parameterize         core/engine.h:31-32,
 *     ExecutableBase::_parameterize   src/engines/base.cpp:9-18     trx_parameterize
 *   Executable::_output               core/engine.h:29-30,
 *     ExecutableBase::_output         src/engines/base.cpp:33-40    trx_output
 *   Executable::inputBufferSize/outputBufferSize  core/engine.h:61-68  trx_program_buffer_bytes
 *
 * The reference executes a tape whose scalar type is Batch<double,8>: eight
 * rollouts ride in the lanes of every slot (dexopt/src/dexopt.cpp:16-25).  This
 * engine generalises "8 lanes" to "R lanes" (R a multiple of the tape's lane
 * width): a trx_memory holds R/8 lane tiles, each tile being exactly one
 * reference memory image, and one launch executes the tape on all tiles.
 *
 * Packed port buffers ("wide" layout).  The reference packs ports in
 * declaration order, each port component-major / lane-minor
 * (src/core/recorder.cpp:514-572, geometry/vector3.h:12-13, core/batch.h:22-29).
 * The wide layout keeps that order with 8 -> R:
 *     scalar-typed port (double):        size bytes, once
 *     lane-typed port (Batch<double,8>): [component][R lanes] doubles
 * For R == 8 the wide layout is byte-identical to the reference's Buffer.
 * Scalar-typed ports are shared by all tiles: inputs of compute/forward
 * programs are broadcast, outputs are read from tile 0; for reverse-mode
 * programs (bprop) scalar-typed input gradients are applied once and
 * scalar-typed output gradients are summed over the tiles -- the reduction the
 * reference performs lane-wise in reverse_batch (dexopt/src/ops.h:22-34).
 *
 * There is no CPU execution path behind this ABI: without a CUDA device every
 * call that touches device state fails with TRX_ERR_CUDA.
 * All functions return TRX_OK (0) or a negative status; trx_last_error() gives
 * the message (thread local), mirroring the reference's std::runtime_error
 * texts (core/engine.h:20-24, src/core/engine.cpp:44-48).
 */
#ifndef TRX_B200_H
#define TRX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int trx_status;
enum {
  TRX_OK = 0,
  TRX_ERR_INVALID = -1,     /* bad argument / malformed program                  */
  TRX_ERR_UNSUPPORTED = -2, /* operator not implemented on the device (no fallback) */
  TRX_ERR_CUDA = -3,        /* CUDA runtime failure or no device                 */
  TRX_ERR_SIZE = -4         /* buffer size mismatch                              */
};

typedef struct trx_program trx_program;
typedef struct trx_memory trx_memory;

/* One operator argument, as Operator::Argument reports it (core/operator.h:84-103). */
typedef struct trx_arg_desc {
  uint32_t size;    /* bytes in the reference memory image */
  uint8_t is_input; /* isInput() */
  uint8_t lanes;    /* 1 = scalar-typed, 8 = Batch<double,8>-typed */
  uint8_t elem;     /* 0 = f64, 1 = u64 */
  uint8_t _pad;
} trx_arg_desc;

typedef struct trx_op_desc {
  const char *name; /* Operator::name(), e.g. "prepare_mul_fg_pose_8d" */
  uint32_t n_args;
  const trx_arg_desc *args;
} trx_op_desc;

/* Program::Input / Output / Parameter / Constant (core/program.h:84-179). */
typedef struct trx_port_desc {
  uint64_t address; /* byte address in the memory image */
  uint64_t offset;  /* byte offset in the reference's packed buffer / const blob */
  uint32_t size;
  uint8_t lanes;
  uint8_t elem;
  uint16_t _pad;
} trx_port_desc;

enum { TRX_SCALAR_AUTO = 0, TRX_SCALAR_SHARED = 1, TRX_SCALAR_REDUCE = 2 };

/* A tractor::Program in neutral form (core/program.h:193-203). */
typedef struct trx_program_desc {
  uint64_t memory_size; /* Program::memorySize() */
  uint32_t n_ops;
  const trx_op_desc *ops; /* operator table */
  uint64_t n_instructions;
  uint64_t n_code_words;
  const uint64_t *code; /* per instruction: op index, then one byte address per argument */
  uint32_t n_inputs, n_parameters, n_outputs, n_constants;
  const trx_port_desc *inputs, *parameters, *outputs, *constants;
  uint64_t n_const_data;
  const uint8_t *const_data;
  /* How scalar-typed ports behave across lane tiles (see header comment).
   * TRX_SCALAR_AUTO derives it from the operator modes on the tape. */
  int32_t scalar_inputs;  /* AUTO | SHARED (broadcast) | REDUCE (apply once)   */
  int32_t scalar_outputs; /* AUTO | SHARED (tile 0)    | REDUCE (sum of tiles) */
} trx_program_desc;

/* ---- library ------------------------------------------------------------ */
const char *trx_last_error(void);
const char *trx_version(void);
/* number of CUDA devices visible; 0 means nothing below can run */
int trx_device_count(void);
/* operator vocabulary: number of device operators and their tape names */
int trx_op_count(void);
const char *trx_op_name(int index);          /* "[mode_]base" without type suffix */
const char *trx_op_signature(int index);     /* e.g. "iT7 iT7 oT7" */

/* ---- programs (Executable) ----------------------------------------------- */
trx_status trx_program_create(const trx_program_desc *desc, int device, trx_program **out);
/* `blob` = one program payload of a TRXB bundle (include/trx_file.h, encodeProgram) */
trx_status trx_program_create_from_blob(const void *blob, size_t bytes, int device, trx_program **out);
void trx_program_destroy(trx_program *p);

enum {
  TRX_BUF_INPUT = 0,
  TRX_BUF_PARAMETER = 1,
  TRX_BUF_OUTPUT = 2
};
/* bytes of the wide packed buffer of `which` ports for `lanes` rollouts */
trx_status trx_program_buffer_bytes(const trx_program *p, int which, uint64_t lanes, uint64_t *bytes);
/* statistics of the lowered program: n_instructions, n_levels, n_bundles, stream_words, ... */
trx_status trx_program_stat(const trx_program *p, const char *key, uint64_t *value);

/* ---- memory (Memory) ------------------------------------------------------ */
trx_status trx_memory_create(uint64_t lanes, int device, trx_memory **out);
void trx_memory_destroy(trx_memory *m);
trx_status trx_memory_lanes(const trx_memory *m, uint64_t *lanes);
/* device bytes currently held (grows on demand like MemoryImpl::resize, engines/base.h:24) */
trx_status trx_memory_bytes(const trx_memory *m, uint64_t *bytes);

/* ---- execution ------------------------------------------------------------ */
/* `stream` is a cudaStream_t (0 = default stream).  Host-pointer variants copy
 * through the memory's pinned staging buffer and are synchronous on return for
 * outputs; *_device variants take device pointers in the wide layout and are
 * fully asynchronous on `stream`. */
trx_status trx_input(const trx_program *p, trx_memory *m, const void *packed, uint64_t bytes, void *stream);
trx_status trx_parameterize(const trx_program *p, trx_memory *m, const void *packed, uint64_t bytes, void *stream);
trx_status trx_execute(const trx_program *p, trx_memory *m, void *stream);
trx_status trx_output(const trx_program *p, trx_memory *m, void *packed, uint64_t bytes, void *stream);

trx_status trx_input_device(const trx_program *p, trx_memory *m, const void *d_packed, uint64_t bytes, void *stream);
trx_status trx_parameterize_device(const trx_program *p, trx_memory *m, const void *d_packed, uint64_t bytes,
                                   void *stream);
trx_status trx_output_device(const trx_program *p, trx_memory *m, void *d_packed, uint64_t bytes, void *stream);

/* In-tape random operators (add_random_normal, add_random_uniform, dropout2: dexopt/src/ops.h:120-204) draw fresh
 * values at every execute, as the reference's do; here every draw is a function of (seed, execute counter, rollout,
 * instruction) -- Philox4x32-10 -- so a run can be repeated: setting the seed restarts the execute counter. */
void trx_set_random_seed(uint64_t seed);

/* Collision shapes for the operators collision_axes and collision_project_2 (dexopt/src/ops.h:208-252,288-375).
 * The reference's tape carries host pointers to CollisionShape / ShapeCollisionPair objects
 * (tractor/include/tractor/collision/shape.h:51-347, collision/query.h:123-152) as uint64 constants and calls Bullet
 * GJK/EPA on the host per lane; here the same constants are handles into a process-wide shape table that the library
 * mirrors to every device it runs on, and the queries run per lane inside the kernel (csrc/collision/trx_collide.h).
 *   kind TRX_SHAPE_SPHERE:     center, radius                          (shape.h:51-63)
 *   kind TRX_SHAPE_CYLINDER:   radius, length, axis z, centred         (shape.h:139-171)
 *   kind TRX_SHAPE_POLYHEDRON: points (support mapping), planes (nx ny nz offset, inside <= 0; project());
 *                              center = mean of the points             (shape.h:196-222,327-346)
 * A pair names the two shapes of one collision_axes instruction.  Handles stay valid until trx_collision_clear;
 * a program created before a clear refuses to execute afterwards (TRX_ERR_INVALID) instead of reading dead records. */
enum { TRX_SHAPE_SPHERE = 1, TRX_SHAPE_CYLINDER = 2, TRX_SHAPE_POLYHEDRON = 3 };
typedef struct trx_shape_desc {
  int kind;
  double center[3]; /* sphere; ignored for polyhedra (computed) and cylinders (0) */
  double radius, length;
  uint32_t n_points, n_planes;
  const double *points; /* 3 * n_points */
  const double *planes; /* 4 * n_planes */
} trx_shape_desc;
trx_status trx_collision_shape_create(const trx_shape_desc *desc, uint64_t *handle);
trx_status trx_collision_pair_create(uint64_t shape_a, uint64_t shape_b, uint64_t *handle);
trx_status trx_collision_shape_center(uint64_t shape, double *center3);
void trx_collision_clear(void);
/* copy of the table (doubles; records as csrc/collision/trx_collide.h lays them out): *count = its length,
 * up to `capacity` doubles are written to `out` (out may be NULL to ask for the length) */
trx_status trx_collision_table(double *out, uint64_t capacity, uint64_t *count);
/* the two queries themselves on the device, one thread per query (tests, and callers outside a tape):
 * poses = n x 14 doubles (pose a, pose b as translation xyz + quaternion xyzw), out = n x 10 (a, b, normal, distance);
 * points = n x 3, out = n x 4 (normal, distance).  Host pointers. */
trx_status trx_collision_query_pairs(uint64_t pair, const double *poses, uint64_t n, double *out, int device);
trx_status trx_collision_project_points(uint64_t shape, const double *points, uint64_t n, double *out, int device);

/* number of kernels this library has launched since load (for launch accounting) */
uint64_t trx_kernel_launch_count(void);

/* development aid: per-level barrier arrival / release clocks of tile 0 (see csrc/trx_engine.cu) */
trx_status trx_debug_trace(uint64_t n_words, long long **device_buffer);
trx_status trx_debug_level_info(const trx_program *p, uint32_t *out, uint64_t n);

/* debug: copy `bytes` of tile `tile`'s memory image starting at byte `address` to the host */
trx_status trx_memory_peek(trx_memory *m, uint64_t tile, uint64_t address, void *dst, uint64_t bytes);

/* ---- objective + gradient (the solver's inner loop) -----------------------------
 * One call = the engine work of one GradientDescentSolver::_step / the first half of
 * LeastSquaresSolver::_step (tractor/include/tractor/solvers/gd.h:50-63, solvers/sq.h:69-77):
 *     x_prog->run(x, mem, r);  loss = r.squaredNorm();  x_prep->execute(mem);  x_bprop->run(r, mem, g);
 * with r kept on the device.  `lanes` rollouts are evaluated; if their write-once memory images
 * exceed `max_device_bytes` (0 = no limit) the rollouts are walked in equal chunks and the
 * gradients summed (rollouts are independent). */
typedef struct trx_objective trx_objective;
/* flags: TRX_OBJECTIVE_FUSED lowers forward+prepare and the adjoint with chain scheduling and on-chip
 * slot rings (csrc/trx_fuse.h) instead of as three generic executables.  The three programs are given
 * as TRXB program payloads (include/trx_file.h); the objective owns everything it builds from them. */
enum {
  TRX_OBJECTIVE_FUSED = 1,
  /* multi-GPU sharding: lane-independent (scalar-typed) residual rows are identical on every rank and must
   * count once over the job; every rank except one sets this flag */
  TRX_OBJECTIVE_NO_SCALAR_ROWS = 2,
  /* lower x_prog, x_prep, x_bprop as three separate executables (default: x_prep is interleaved into
   * x_prog and the two sweeps are two launches) */
  TRX_OBJECTIVE_SEPARATE = 4,
  /* keep every operand copy of x_prep (default: where a prepare rule only copies an operand, the adjoint reads the
   * forward slot instead and the copy is dropped, csrc/trx_direct.h).  Needed when a tangent program is attached
   * (trx_objective_attach_fprop): the forward_* rules read the saved operands as well. */
  TRX_OBJECTIVE_KEEP_PREP_STATE = 8,
  /* re-rolled objectives: what a frame's checkpoint holds beyond the carried state.  Default: everything the
   * adjoint reads (link poses, policy values, contact points), so the adjoint sweep recomputes nothing.
   * _POLICY: the policy layers' values only (kinematics and contacts are recomputed); _STATE: the carried state
   * only (the whole frame is recomputed, least HBM traffic).  Results are identical; DESIGN.md section 3 has the
   * traffic / time of each. */
  TRX_OBJECTIVE_CHECKPOINT_STATE = 16,
  TRX_OBJECTIVE_CHECKPOINT_POLICY = 32
};
trx_status trx_objective_create(const void *prog_blob, size_t prog_bytes, const void *prep_blob, size_t prep_bytes,
                                const void *bprop_blob, size_t bprop_bytes, uint64_t lanes, uint64_t max_device_bytes,
                                int device, int flags, trx_objective **out);
/* The same objective in re-rolled form (csrc/rollout/rollout_plan.h): where
 * DexLearn::_runBatch (dexopt/src/dexlearn.h:393-422) unrolls its frame loop onto one flat tape, the product's
 * assembly can also emit ONE frame program plus the state carried between frames; it is evaluated by the
 * hand-written forward / adjoint frame kernels of csrc/trx_rollout.cu (forward kinematics, contacts and the
 * adjoint fused, policy layers on the FP64 tensor cores, one state checkpoint per frame in HBM).  `options` is
 * the problem description trx_dexsyn_build takes; results equal those of trx_objective_create over the tapes of
 * trx_dexsyn_build(options).  TRX_ERR_UNSUPPORTED when the problem has no re-rolled form (dropout masks,
 * outer > 1, no policy): use the tape objective.  Every other trx_objective_* call works on it. */
trx_status trx_objective_create_rollout(const char *options, uint64_t lanes, uint64_t max_device_bytes, int device,
                                        int flags, trx_objective **out);
/* r = x_prog->run(x) (tractor/include/tractor/solvers/gd.h:50-53): the residual vector of all rollouts in the
 * wide layout (lane-independent rows once, then [row][rollout]); info key "residual_elems" gives its length.
 * Re-rolled objectives only (tape objectives: trx_execute + trx_output on the prog executable). */
trx_status trx_objective_residuals(trx_objective *o, const double *x, double *r, uint64_t r_bytes, void *stream);
void trx_objective_destroy(trx_objective *o);
/* keys: n_x lanes chunk_lanes n_chunks parameter_bytes device_bytes launches_per_evaluation fused,
 * and prog.<stat> / prep.<stat> / bprop.<stat> with the keys of trx_program_stat */
trx_status trx_objective_info(const trx_objective *o, const char *key, uint64_t *value);
/* per-lane parameters of all rollouts, wide layout, host pointer (Executable::parameterize) */
trx_status trx_objective_parameterize(trx_objective *o, const void *packed, uint64_t bytes, void *stream);
/* host buffers: x[n_x] in, *loss and gradient[n_x] out; H2D and D2H copies included; synchronous */
trx_status trx_objective_evaluate(trx_objective *o, const double *x, double *loss, double *gradient, void *stream);
/* device buffers: d_x[n_x] in, d_out[1 + n_x] = {loss, gradient} out; asynchronous on `stream` */
trx_status trx_objective_evaluate_device(trx_objective *o, const double *d_x, double *d_out, void *stream);
/* GradientDescentSolver::_step (tractor/include/tractor/solvers/gd.h:43-129) n_steps times without leaving
 * the device: engine work as trx_objective_evaluate_device, then v <- momentum v - learning_rate g ; x <- x + v
 * (gd.h:74,96; skipped when a gradient element is not finite, gd.h:72,124-126).  d_x[n_x] in/out,
 * d_v[n_x] in/out (zero it for a fresh solver), d_losses[n_steps] out or NULL: the loss before each update
 * (what Solver::loss() returns after step(), solver.cpp:96-103).  Asynchronous on `stream`. */
trx_status trx_objective_gd_steps(trx_objective *o, double *d_x, double *d_v, double learning_rate, double momentum,
                                  uint32_t n_steps, double *d_losses, void *stream);
/* ---- Gauss-Newton (LeastSquaresSolver, tractor/include/tractor/solvers/sq.h:57-123) -------------------------
 * The tangent program x_fprop of the same bundle (buildGradients, gradients.cpp:236-316) is attached to a tape
 * objective; J^T J v is then x_fprop ; x_bprop on the objective's memory, which is what the reference's x_hprop
 * executes (gradients.cpp:344-362) and what MatrixReplacement::mul hands to Eigen's conjugate gradients
 * (core/eigen.h:73-77).  Lane-independent residual rows seed the adjoint once, scalar gradients are summed over
 * lane tiles and, with a communicator, over ranks.  Tape objectives with all rollouts resident only. */
trx_status trx_objective_attach_fprop(trx_objective *o, const void *fprop_blob, size_t fprop_bytes);
/* d_hv[n_x] = J^T J d_v at the point of the last evaluation; device pointers, asynchronous on `stream` */
trx_status trx_objective_hessian_product(trx_objective *o, const double *d_v, double *d_hv, void *stream);
/* n_steps of LeastSquaresSolver::_step with x, g, the linear solution and the conjugate-gradient vectors resident
 * on the device: solve (J^T J + regularization I) d = J^T r by conjugate gradients (Eigen::ConjugateGradient with
 * the identity preconditioner: zero initial guess, stop at |residual|^2 < tolerance^2 |J^T r|^2 or after
 * max_linear_iterations, 0 = n_x), then x <- x - step_scaling d.  dexopt: 0.1, 100, 0.5, 1e-9
 * (dexopt/src/dexopt.cpp:153-160).  d_x[n_x] in/out (device); losses[n_steps] and linear_iterations[n_steps] are
 * host arrays or NULL.  Synchronous. */
trx_status trx_objective_sq_steps(trx_objective *o, double *d_x, double regularization, uint32_t max_linear_iterations,
                                  double step_scaling, double tolerance, uint32_t n_steps, double *losses,
                                  uint32_t *linear_iterations, void *stream);
/* CUDA-event times of the last trx_objective_evaluate: {h2d, kernels, d2h} in milliseconds */
trx_status trx_objective_last_times(const trx_objective *o, float *ms3);
/* per-program kernel timing with CUDA events on the launching stream: enable, then call
 * trx_objective_profile_collect after each evaluation; ms3 = running totals {prog, prep+seed, bprop} */
trx_status trx_objective_set_profiling(trx_objective *o, int enable);
trx_status trx_objective_profile_collect(trx_objective *o, double *ms3, uint64_t *evaluations);

/* ---- multi-GPU: rollouts sharded over the ranks of one job ---------------------------------
 * The reference has no multi-device path; its "outer batch" is a sum of independent tape replicas
 * (dexopt/src/dexlearn.h:440-453) and the only cross-lane operation is the sum behind reverse_batch
 * (dexopt/src/ops.h:22-34).  Here every rank owns an objective over its share of the rollouts (all ranks but one
 * with TRX_OBJECTIVE_NO_SCALAR_ROWS) and a communicator; with a communicator attached, every evaluation
 * (trx_objective_evaluate[_device], each step of trx_objective_gd_steps, trx_objective_hessian_product and the
 * Gauss-Newton steps built on it) ends with ONE ncclAllReduce(sum) of {loss, gradient[n_x]} on the caller's
 * stream, so all ranks hold the job's loss and gradient and the device-resident loops need no host in between.
 * NCCL is bound at run time (libnccl.so.2); without it trx_comm_* fail with TRX_ERR_UNSUPPORTED.
 *   rank 0:  trx_comm_unique_id(id)  -> broadcast the 128 bytes by any means (MPI, torch.distributed, a file)
 *   all:     trx_comm_create(id, n_ranks, rank, device, &comm);  trx_objective_set_comm(objective, comm) */
typedef struct trx_comm trx_comm;
enum { TRX_COMM_ID_BYTES = 128 };
trx_status trx_comm_unique_id(uint8_t *id /* TRX_COMM_ID_BYTES */);
trx_status trx_comm_create(const uint8_t *id, int n_ranks, int rank, int device, trx_comm **out);
void trx_comm_destroy(trx_comm *c);
/* sum `count` doubles in place over the ranks (what the objective calls do with {loss, gradient}) */
trx_status trx_comm_allreduce(trx_comm *c, double *d_values, uint64_t count, void *stream);
/* the objective does not own the communicator; NULL detaches */
trx_status trx_objective_set_comm(trx_objective *o, trx_comm *comm);

/* ---- objective assembly (product-side recorder) ------------------------------
 * Stands where dexopt records its problem onto a tape: DexLearn::build ->
 * Program([&]{ _runTraining(); }) (dexopt/src/dexlearn.h:440-453,511-518) followed by
 * SolverBase::_compileGradients -> buildGradients (tractor/include/tractor/solvers/base.h:112-138,
 * tractor/src/core/gradients.cpp:22-407).  Builds the synthetic Shadow-hand rollout problem
 * (csrc/assembly) with the product's own tape recorder and differentiates it.
 * `options`: space separated key=value pairs: kind=handarm|hand8|chain frames=T outer=N hidden=H
 * dropout=0|1 seed=N joints=N contacts=N extra_fixed=N wstd=S.
 * On success *bundle is a malloc'ed TRXB bundle (include/trx_file.h) holding programs
 * prog/prep/fprop/bprop/hprop/accu, the array "x" (initial policy parameters) and text "meta". */
int trx_dexsyn_build(const char *options, uint8_t **bundle, uint64_t *bundle_bytes);
void trx_builder_free(uint8_t *bundle);
const char *trx_builder_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* TRX_B200_H */
