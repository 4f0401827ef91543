"""Host mirror of the optimiser loop that calls the hot path.

Reference being mirrored:
    tractor/include/tractor/solvers/base.h:112-182   SolverBase: compiles prog/prep/bprop/... on one Memory
    tractor/include/tractor/solvers/gd.h:43-129      GradientDescentSolver::_step
    tractor/src/core/solver.cpp:52-103               Solver::gather / solve / scatter / step / loss
The engine work of one step (x_prog, loss, x_prep, x_bprop) is a single C-ABI call,
trx_objective_evaluate (include/trx_b200.h); the update rule runs on the host on the
8 083-element parameter vector exactly as gd.h:74,96 write it:  v <- mu v - lr g ;  x <- x (+) v.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _native, trxfile
from ._native import check
from .engine import CudaEngine


class Objective:
    """objective + gradient over R rollouts of a recorded problem (prog/prep/bprop of one bundle)"""

    def __init__(self, engine: CudaEngine, bundle: trxfile.Bundle, lanes: int, max_device_bytes: int = 0,
                 fused: bool = False, scalar_rows: bool = True, separate: bool = False, gauss_newton: bool = False):
        """default: x_prep interleaved into x_prog, level-scheduled (csrc/trx_lower.h), two launches per chunk.
        separate=True: prog, prep, bprop as three generic executables.  fused=True: experimental chain
        scheduling with on-chip slot rings (csrc/trx_fuse.h).
        scalar_rows=False: this rank does not count the lane-independent residual rows (multi-GPU).
        gauss_newton=True: keep x_prep's operand copies so that a tangent program can be attached (attach_fprop)"""
        self.engine = engine
        self._h = ctypes.c_void_p()
        blobs = [bundle.programs[n] for n in ("prog", "prep", "bprop")]
        bufs = [ctypes.create_string_buffer(b, len(b)) for b in blobs]
        check(_native.lib().trx_objective_create(bufs[0], len(blobs[0]), bufs[1], len(blobs[1]), bufs[2], len(blobs[2]),
                                                 lanes, max_device_bytes, engine.device,
                                                 (1 if fused else 0) | (0 if scalar_rows else 2) | (4 if separate else 0) |
                                                 (8 if gauss_newton else 0),
                                                 ctypes.byref(self._h)))
        self.lanes = lanes
        self.n_x = self.info("n_x")
        self._g = np.empty(self.n_x)

    def info(self, key: str) -> int:
        v = ctypes.c_uint64()
        check(_native.lib().trx_objective_info(self._h, key.encode(), ctypes.byref(v)))
        return v.value

    def parameterize(self, wide_params: np.ndarray) -> None:
        b = np.ascontiguousarray(wide_params, dtype=np.float64)
        check(_native.lib().trx_objective_parameterize(self._h, b.ctypes.data, b.nbytes, None))

    def set_comm(self, comm) -> None:
        """rollouts sharded over ranks (distributed.Communicator): from now on every evaluation -- also inside the
        device-resident solver loops -- ends with one all-reduce of {loss, gradient} on its stream"""
        self._comm = comm  # keep it alive
        check(_native.lib().trx_objective_set_comm(self._h, comm._h if comm is not None else None))

    def evaluate(self, x: np.ndarray):
        """host x -> (loss, gradient); H2D of x and D2H of (loss, g) happen inside"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n_x
        loss = ctypes.c_double()
        check(_native.lib().trx_objective_evaluate(self._h, x.ctypes.data, ctypes.byref(loss), self._g.ctypes.data,
                                                   None))
        return loss.value, self._g.copy()

    def evaluate_device(self, d_x_ptr: int, d_out_ptr: int, stream: int = 0) -> None:
        check(_native.lib().trx_objective_evaluate_device(self._h, ctypes.c_void_p(d_x_ptr),
                                                          ctypes.c_void_p(d_out_ptr), ctypes.c_void_p(stream)))

    def gd_steps(self, d_x_ptr: int, d_v_ptr: int, learning_rate: float, momentum: float, n_steps: int,
                 d_losses_ptr: int = 0, stream: int = 0) -> None:
        """n_steps GradientDescentSolver steps on the device (no host round trip between them)"""
        check(_native.lib().trx_objective_gd_steps(self._h, ctypes.c_void_p(d_x_ptr), ctypes.c_void_p(d_v_ptr),
                                                   ctypes.c_double(learning_rate), ctypes.c_double(momentum),
                                                   ctypes.c_uint32(n_steps), ctypes.c_void_p(d_losses_ptr),
                                                   ctypes.c_void_p(stream)))

    def attach_fprop(self, bundle: trxfile.Bundle) -> None:
        """the tangent program of the same bundle: enables hessian_product and LeastSquaresSolver (tape objectives)"""
        blob = bundle.programs["fprop"]
        buf = ctypes.create_string_buffer(blob, len(blob))
        check(_native.lib().trx_objective_attach_fprop(self._h, buf, len(blob)))

    def hessian_product(self, v: np.ndarray) -> np.ndarray:
        """J^T J v at the point of the last evaluate() -- the reference's x_hprop (core/eigen.h:73-77)"""
        import torch
        dev = torch.device("cuda", self.engine.device)
        d_v = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
        d_hv = torch.empty(self.n_x, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(_native.lib().trx_objective_hessian_product(self._h, ctypes.c_void_p(d_v.data_ptr()),
                                                          ctypes.c_void_p(d_hv.data_ptr()), ctypes.c_void_p(stream)))
        return d_hv.cpu().numpy()

    def sq_steps(self, d_x_ptr: int, regularization: float, max_linear_iterations: int, step_scaling: float,
                 tolerance: float, n_steps: int, stream: int = 0):
        """n_steps LeastSquaresSolver steps with all vectors on the device; returns (losses, linear iterations)"""
        losses = np.zeros(max(1, n_steps))
        its = np.zeros(max(1, n_steps), dtype=np.uint32)
        check(_native.lib().trx_objective_sq_steps(self._h, ctypes.c_void_p(d_x_ptr), ctypes.c_double(regularization),
                                                   ctypes.c_uint32(max_linear_iterations), ctypes.c_double(step_scaling),
                                                   ctypes.c_double(tolerance), ctypes.c_uint32(n_steps),
                                                   losses.ctypes.data, its.ctypes.data, ctypes.c_void_p(stream)))
        return losses[:n_steps], its[:n_steps]

    def set_profiling(self, enable: bool) -> None:
        check(_native.lib().trx_objective_set_profiling(self._h, 1 if enable else 0))

    def profile_collect(self):
        """running totals of CUDA-event kernel time {prog, prep+seed, bprop} in ms, and evaluation count"""
        ms = (ctypes.c_double * 3)()
        n = ctypes.c_uint64()
        check(_native.lib().trx_objective_profile_collect(self._h, ms, ctypes.byref(n)))
        return [float(ms[0]), float(ms[1]), float(ms[2])], n.value

    def last_times_ms(self):
        ms = (ctypes.c_float * 3)()
        check(_native.lib().trx_objective_last_times(self._h, ms))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _native.lib().trx_objective_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


class RolloutObjective(Objective):
    """The same objective in re-rolled form (trx_objective_create_rollout): one frame program and the carried
    state instead of a flat tape, evaluated by the hand-written forward / adjoint frame kernels
    (csrc/trx_rollout.cu).  Built from the problem description that build_dexsyn takes; per-rollout
    parameters are [3][lanes] = object start offset x, y and yaw (dexenv_grasp.h:149-165)."""

    #: what a frame's checkpoint holds (include/trx_b200.h TRX_OBJECTIVE_CHECKPOINT_*): "full" = everything the
    #: adjoint reads (default), "policy" = carried state + policy values, "state" = the carried state only
    CHECKPOINT_FLAGS = {"full": 0, "policy": 32, "state": 16}

    def __init__(self, engine: CudaEngine, lanes: int, max_device_bytes: int = 0, scalar_rows: bool = True,
                 checkpoint: str = "full", **options):
        self.engine = engine
        self._h = ctypes.c_void_p()
        self.options = " ".join("%s=%s" % (k, v) for k, v in options.items())
        L = _native.lib()
        L.trx_objective_create_rollout.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                                   ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.trx_objective_residuals.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                              ctypes.c_void_p]
        check(L.trx_objective_create_rollout(self.options.encode(), lanes, max_device_bytes, engine.device,
                                             (0 if scalar_rows else 2) | self.CHECKPOINT_FLAGS[checkpoint],
                                             ctypes.byref(self._h)))
        self.lanes = lanes
        self.n_x = self.info("n_x")
        self._g = np.empty(self.n_x)

    def residuals(self, x: np.ndarray) -> np.ndarray:
        """r = x_prog->run(x): lane-independent rows once, then [row][rollout]"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n_x
        r = np.empty(self.info("residual_elems"))
        check(_native.lib().trx_objective_residuals(self._h, x.ctypes.data, r.ctypes.data, r.nbytes, None))
        return r


class GradientDescentSolver:
    """GradientDescentSolver (solvers/gd.h:14-129): momentum gradient descent on J^T r."""

    def __init__(self, objective: Objective, learning_rate: float = 0.01, momentum: float = 0.0):
        self.objective = objective
        self.learning_rate = learning_rate  # dexopt.cpp:199-202 uses 0.01 / 0
        self.momentum = momentum
        self.x: Optional[np.ndarray] = None
        self.v: Optional[np.ndarray] = None
        self._loss = float("nan")

    def gather(self, x0: np.ndarray) -> None:  # solver.cpp:52-60
        self.x = np.array(x0, dtype=np.float64)
        self.v = np.zeros_like(self.x)

    def step(self) -> float:  # gd.h:43-129
        loss, g = self.objective.evaluate(self.x)
        self._loss = loss
        if np.all(np.isfinite(g)):  # gd.h:72,124-126: non-finite gradients skip the update
            self.v = self.momentum * self.v - self.learning_rate * g
            self.x = self.x + self.v  # accu program for scalar inputs: plain add (gradients.cpp:396-403)
        return loss

    def steps_on_device(self, n_steps: int) -> np.ndarray:
        """n_steps steps with x, the momentum state and the update rule resident on the GPU
        (trx_objective_gd_steps): one upload of x, one download of {x, losses}.  Returns the losses."""
        import torch
        dev = torch.device("cuda", self.objective.engine.device)
        d_x = torch.from_numpy(self.x).to(dev)
        d_v = torch.from_numpy(self.v).to(dev)
        d_losses = torch.empty(max(1, n_steps), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.objective.gd_steps(d_x.data_ptr(), d_v.data_ptr(), self.learning_rate, self.momentum, n_steps,
                                d_losses.data_ptr(), stream)
        self.x = d_x.cpu().numpy()
        self.v = d_v.cpu().numpy()
        losses = d_losses.cpu().numpy()[:n_steps]
        if n_steps:
            self._loss = float(losses[-1])
        return losses

    def loss(self) -> float:  # gd.h:40-41
        return self._loss

    def scatter(self) -> np.ndarray:  # solver.cpp:62-66
        return self.x.copy()


class LeastSquaresSolver:
    """LeastSquaresSolver (solvers/sq.h:14-123): Gauss-Newton steps, (J^T J + lambda I) d = J^T r solved by conjugate
    gradients over the Gauss-Newton product (core/eigen.h:53-83, solvers/base.h:76-85), x <- x - step_scaling d.
    Defaults are dexopt's "sq" solver (dexopt/src/dexopt.cpp:153-160).  The objective must be a tape objective with
    its tangent program attached (Objective.attach_fprop); everything runs on the device
    (trx_objective_sq_steps)."""

    def __init__(self, objective: Objective, regularization: float = 0.1, max_linear_iterations: int = 100,
                 step_scaling: float = 0.5, tolerance: float = 1e-9):
        self.objective = objective
        self.regularization = regularization
        self.max_linear_iterations = max_linear_iterations
        self.step_scaling = step_scaling
        self.tolerance = tolerance
        self.x: Optional[np.ndarray] = None
        self._loss = float("nan")
        self.linear_iterations = []

    def gather(self, x0: np.ndarray) -> None:
        self.x = np.array(x0, dtype=np.float64)

    def steps(self, n_steps: int) -> np.ndarray:
        import torch
        dev = torch.device("cuda", self.objective.engine.device)
        d_x = torch.from_numpy(self.x).to(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        losses, its = self.objective.sq_steps(d_x.data_ptr(), self.regularization, self.max_linear_iterations,
                                              self.step_scaling, self.tolerance, n_steps, stream)
        self.x = d_x.cpu().numpy()
        self.linear_iterations.extend(int(i) for i in its)
        if n_steps:
            self._loss = float(losses[-1])
        return losses

    def step(self) -> float:
        return float(self.steps(1)[0])

    def loss(self) -> float:
        return self._loss

    def scatter(self) -> np.ndarray:
        return self.x.copy()
