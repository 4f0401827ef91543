#!/usr/bin/env python
"""bench.py -- objective+gradient throughput of dexopt's recorded-objective hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): objective+gradient evaluations per second, counted in timestep*rollouts/s.
One step = one evaluation of the objective and its reverse-mode gradient
(x_prog -> loss -> x_prep -> x_bprop, tractor/include/tractor/solvers/gd.h:50-63) over all T frames of
all R rollouts of the synthetic "Shadow hand + UR arm" problem of BASELINE.json configs[2]:
29 controlled joints, policy MLP 41->64->83, horizon 256, 1024 rollouts per GPU (weak scaling:
configs[3] is 8192 rollouts over 8 GPUs, with one NCCL all-reduce of {loss, gradient} per step;
BENCH_TOTAL_ROLLOUTS=8192 fixes the total instead: strong scaling).

Measured arm: the re-rolled rollout objective (trx_objective_create_rollout: hand-written forward / adjoint
frame kernels, csrc/trx_rollout.cu).  BENCH_ENGINE=tape measures the interpreted flat tape instead
(trx_objective_create over the tapes of trx_dexsyn_build).  Nothing under oracle/ is used by the measured arm;
`--impl reference`, the `cpu_baseline` leg and the parity check before the timed region run the reference's own
SimpleEngine (oracle/_ref, built from the reference sources) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES = int(os.environ.get("BENCH_FRAMES", 256))
LANES_PER_GPU = int(os.environ.get("BENCH_LANES", 1024))
TOTAL_ROLLOUTS = int(os.environ.get("BENCH_TOTAL_ROLLOUTS", 0))  # > 0: strong scaling, rollouts split over the ranks
KIND = os.environ.get("BENCH_KIND", "handarm")
HIDDEN = int(os.environ.get("BENCH_HIDDEN", 64))
# initial policy weights N(0, WSTD^2): DenseLayer's default stdev (dexopt/src/neural.h:134-136).  (With 0.1 the
# 256-frame rollout leaves every physical regime and its gradient overflows fp64 -- same time per evaluation, no meaning.)
WSTD = float(os.environ.get("BENCH_WSTD", 0.001))
ENGINE = os.environ.get("BENCH_ENGINE", "rollout")  # rollout | tape
FUSE_DENSE = int(os.environ.get("BENCH_FUSE_DENSE", 1))  # tape engine: dense layers as block instructions
MAX_DEVICE_BYTES = int(float(os.environ.get("BENCH_MAX_DEVICE_GB", 150)) * 1e9)
CPU_SAMPLE_FRAMES = int(os.environ.get("BENCH_CPU_FRAMES", 16))
CPU_SAMPLE_SECONDS = float(os.environ.get("BENCH_CPU_SECONDS", 10))
METRIC = "objective+gradient evals/sec (timestep*rollouts/s)"
UNIT = "timestep*rollouts/s"
# ncu --set full captures of the dominant launch (profiles/README.md): dram__bytes_read.sum + dram__bytes_write.sum
# per timestep*rollout.  rollout: profiles/ncu_full_r2_rollout_T256_raw.csv; tape: profiles/ncu_full_r1_staged_T32_raw.csv
NCU_DRAM_BYTES_PER_UNIT = {"rollout": None, "tape": (2.465697e9 + 1.354105e9) / (32 * 1024)}
NCU_SOURCE = {"rollout": "profiles/ncu_full_r2_rollout_T256_raw.csv", "tape": "profiles/ncu_full_r1_staged_T32_raw.csv"}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_r2_rollout_dram.json")) as _f:
        NCU_DRAM_BYTES_PER_UNIT["rollout"] = float(json.load(_f)["dram_bytes_per_unit"])
except Exception:
    pass


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_peak_fp64():
    """DFMA throughput measured on this pool's B200 by scripts/microbench/fp64_peak.cu (MEASURED_PEAKS.json has
    no FP64 figure)"""
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peak_r2.json")) as f:
            d = json.load(f)
            return float(d["fp64_dfma_tflops"]), "measured (profiles/fp64_peak_r2.json: DFMA %.1f, DMMA %.1f TFLOP/s)" % (
                d["fp64_dfma_tflops"], d["fp64_dmma_tflops"])
    except Exception:
        return 40.0, "nominal (no measurement found)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.sm_max = float(parts[1])
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop_flag.wait(0.05)  # a 20-step run of the rollout kernel is ~0.35 s over the three timed arms

    def stop(self):
        self._stop_flag.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_tool")


def reference_cpu(threads, frames, seconds, repeats=1, engine="simple"):
    """the reference's own SimpleEngine (oracle/_ref/libtractor_ref.so) on `threads` host threads, each with
    its own Memory over one recording of the same problem; a bounded sample of the workload.  The clock starts
    when every thread has built its runner; the median of `repeats` windows of `seconds` is reported."""
    if not os.path.exists(REF_TOOL):
        return None
    out = subprocess.run([REF_TOOL, "bench", "kind=" + KIND, "hidden=%d" % HIDDEN, "wstd=%g" % WSTD, "frames=%d" % frames,
                          "threads=%d" % threads, "seconds=%g" % seconds, "repeats=%d" % repeats, "engine=" + engine],
                         capture_output=True, text=True)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def lanes_per_gpu(world):
    if TOTAL_ROLLOUTS:
        assert TOTAL_ROLLOUTS % world == 0
        return TOTAL_ROLLOUTS // world
    return LANES_PER_GPU


def workload_config(n_gpus):
    lanes = lanes_per_gpu(n_gpus)
    cfg = {
        "workload": "BASELINE.json configs[2]%s: synthetic Shadow hand + UR arm (29 controlled joints, 6 end "
                    "effectors, floating object), policy MLP 41->%d->83, horizon %d, %d rollouts per GPU" % (
                        " sharded as configs[3]" if n_gpus > 1 else "", HIDDEN, FRAMES, lanes),
        "frames": FRAMES, "rollouts_per_gpu": lanes, "rollouts": lanes * n_gpus,
        "parallelism": "rollouts sharded over %d GPU(s); one NCCL all-reduce of {loss, gradient} per step, issued by the "
                       "library on the evaluation's stream" % n_gpus if n_gpus > 1
        else "single GPU",
    }
    if ENGINE == "rollout":
        cfg["engine"] = "re-rolled rollout objective (one frame program + carried state; hand-written forward/adjoint kernels)"
        cfg["l2"] = ("every evaluation rewrites its per-frame checkpoints (several hundred MB per GPU, see device_bytes; "
                     "larger than the 126 MB L2) before the adjoint reads them back in reverse order; no flush needed")
    else:
        cfg["engine"] = "interpreted flat tape"
        cfg["dense_layers"] = "one block instruction per layer (x_dense)" if FUSE_DENSE else "reference expansion (batch + dot4add per weight)"
        cfg["l2"] = "working set (write-once tape image, tens of GB per chunk) is far larger than the 126 MB L2; no flush needed"
    return cfg


def run_reference(args, rank, world):
    """--impl reference: the reference SimpleEngine on all host threads; each step is one bounded window"""
    if rank != 0:
        return 0
    threads = host_threads()
    t0 = time.time()
    values = []
    total = args.warmup + args.steps
    per = max(3.0, min(12.0, 240.0 / max(1, total)))
    for i in range(total):
        r = reference_cpu(threads, CPU_SAMPLE_FRAMES, per)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_tool is not built"}))
            return 0
        if i >= args.warmup:
            values.append(r["step_rollouts_per_s"])
    values.sort()
    value = values[len(values) // 2]
    lanes = lanes_per_gpu(args.gpus)
    sample = ("reference SimpleEngine, Batch<double,8> lanes, %d frames of the %d-frame rollout per evaluation (the rate "
              "per timestep*rollout does not depend on the horizon: profiles/ref_cpu_linearity_r2.jsonl), %d threads x own "
              "Memory, %.0f s per step after all threads are ready, median of the %d steps" % (
                  CPU_SAMPLE_FRAMES, FRAMES, threads, per, len(values)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * FRAMES * lanes * args.gpus / value,
        "higher_is_better": True, "scaling": "strong" if TOTAL_ROLLOUTS else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


def parity_check(pkg, engine_obj, np):
    """before anything is timed: loss and gradient of a 16-frame, 16-rollout slice of the benchmarked problem
    against the reference SimpleEngine run right here (ref_tool record); 1e-9 relative as the north star asks"""
    if not os.path.exists(REF_TOOL):
        return {"skipped": "oracle/_ref/ref_tool is not built"}
    from robio_2021_dexopt_b200 import trxfile
    frames = 16
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "slice.trxb")
        subprocess.run([REF_TOOL, "record", path, "problem=dexsyn", "kind=" + KIND, "hidden=%d" % HIDDEN, "wstd=%g" % WSTD,
                        "frames=%d" % frames, "tiles=2", "programs=0", "keep_dr=0"], check=True, capture_output=True)
        ref = trxfile.load(path)
    params = np.concatenate([ref.arrays["tile%d/params" % t].reshape(-1, 8) for t in range(2)], axis=1).copy()
    x = ref.arrays["x"]
    if ENGINE == "rollout":
        obj = pkg.RolloutObjective(engine_obj, 16, kind=KIND, hidden=HIDDEN, frames=frames)
    else:
        bundle = pkg.build_dexsyn(kind=KIND, frames=frames, hidden=HIDDEN, programs="prog,prep,bprop", fuse_dense=FUSE_DENSE)
        obj = pkg.Objective(engine_obj, bundle, lanes=16, separate=True)
    obj.parameterize(params.reshape(-1))
    loss, g = obj.evaluate(x)
    want_loss, want_g = float(ref.arrays["loss_total"][0]), ref.arrays["g_total"]
    loss_err = abs(loss - want_loss) / abs(want_loss)
    g_err = float(np.max(np.abs(g - want_g)) / np.max(np.abs(want_g)))
    assert loss_err <= 1e-9 and g_err <= 1e-9, "parity against the reference failed: loss %.3g gradient %.3g" % (loss_err, g_err)
    return {"slice": "%d frames x 16 rollouts of the benchmarked problem" % frames, "loss_rel_err": loss_err,
            "gradient_max_err_over_max": g_err, "reference_loss": want_loss, "tolerance": 1e-9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=int(os.environ.get("BENCH_STEPS", 10)))
    ap.add_argument("--warmup", type=int, default=int(os.environ.get("BENCH_WARMUP", 3)))
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import robio_2021_dexopt_b200 as pkg
    from robio_2021_dexopt_b200 import flops

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lanes = lanes_per_gpu(world)
    engine = pkg.CudaEngine(local_rank)
    skip_cpu = os.environ.get("BENCH_SKIP_CPU", "0") == "1"
    parity = None
    if rank == 0 and world == 1 and not skip_cpu:
        parity = parity_check(pkg, engine, np)

    t_build = time.time()
    problem = dict(kind=KIND, hidden=HIDDEN, wstd=WSTD)
    if ENGINE == "rollout":
        obj = pkg.RolloutObjective(engine, lanes, max_device_bytes=MAX_DEVICE_BYTES, scalar_rows=(rank == 0),
                                   frames=FRAMES, **problem)
        x_host = pkg.build_dexsyn(frames=1, programs="prog", **problem).arrays["x"].copy()
        n_ports = 3
    else:
        bundle = pkg.build_dexsyn(frames=FRAMES, programs="prog,prep,bprop", fuse_dense=FUSE_DENSE, **problem)
        obj = pkg.Objective(engine, bundle, lanes=lanes, max_device_bytes=MAX_DEVICE_BYTES,
                            separate=bool(int(os.environ.get("BENCH_SEPARATE", 1))), scalar_rows=(rank == 0))
        x_host = bundle.arrays["x"].copy()
        n_ports = len(bundle.view("prog").parameters)
    t_build = time.time() - t_build
    n_x = obj.n_x
    # per-rollout random draws: object start offset x,y ~ U(-0.01, 0.01), yaw ~ U(0, 2 pi)  (dexenv_grasp.h:151-164)
    rng = np.random.default_rng(20211227 + 3 + rank)
    params = np.zeros((n_ports, lanes))
    params[0] = rng.uniform(-0.01, 0.01, lanes)
    params[1] = rng.uniform(-0.01, 0.01, lanes)
    params[2] = rng.uniform(0.0, 2 * np.pi, lanes)
    obj.parameterize(params.reshape(-1))
    units = FRAMES * lanes * world  # timestep*rollouts per step, whole job

    stream = torch.cuda.current_stream()
    d_x = torch.from_numpy(x_host).cuda()
    d_out = torch.zeros(n_x + 1, dtype=torch.float64, device="cuda")
    h_x = torch.from_numpy(x_host.copy()).pin_memory()
    h_out = torch.zeros(n_x + 1, dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        # rollouts sharded over the ranks: the library sums {loss, gradient} over them (one ncclAllReduce on the
        # evaluation's stream, include/trx_b200.h trx_objective_set_comm); torch.distributed only carries the id
        from robio_2021_dexopt_b200.distributed import Communicator
        comm = Communicator.from_torch_distributed(local_rank)
        obj.set_comm(comm)

    def step_device():
        obj.evaluate_device(d_x.data_ptr(), d_out.data_ptr(), stream.cuda_stream)

    def step_e2e():
        # the reference-facing C-ABI call with host buffers (trx_objective_evaluate); with several ranks it ends
        # with the all-reduce, so every rank returns the job's loss and gradient
        return obj.evaluate(x_host)

    # ---- device-resident arm: warm-up, then exactly K timed steps
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = pkg.kernel_launch_count()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = pkg.kernel_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)

    # ---- the engine's launches alone (no all-reduce): CUDA events on the launching stream
    kernel_ms = None
    tape_ms = None
    if world > 1:
        obj.set_comm(None)
    if ENGINE == "rollout":
        k0 = torch.cuda.Event(enable_timing=True)
        k1 = torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(args.steps):
            obj.evaluate_device(d_x.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        k1.record()
        torch.cuda.synchronize()
        kernel_ms = k0.elapsed_time(k1) / args.steps
    else:
        obj.set_profiling(True)
        for _ in range(args.steps):
            obj.evaluate_device(d_x.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
            tape_ms, _n = obj.profile_collect()
        obj.set_profiling(False)

    if world > 1:
        obj.set_comm(comm)
    # ---- end-to-end arm: host buffers in, host (loss, gradient) out, every step
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss, g = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # the objective the timed steps evaluated is a meaningful one: finite loss and gradient
    assert np.isfinite(loss) and np.all(np.isfinite(g)), "the benchmarked objective is not finite"
    ms_per_step = ms_total / args.steps
    value = units / (ms_per_step * 1e-3)
    e2e_value = units / (e2e_s / args.steps)
    peak, peak_source = measured_peak_hbm()
    n_chunks = obj.info("n_chunks")
    units_per_launch = FRAMES * obj.info("chunk_lanes")
    n_q = 29 if KIND == "handarm" else 24 if KIND == "hand8" else int(os.environ.get("BENCH_JOINTS", 30))
    b_alg = 2 * 8 * (n_q + 13)  # state checkpoint written by the forward sweep, read by the adjoint
    if ENGINE == "rollout":
        launch_ms = kernel_ms / n_chunks
        alg_bytes = b_alg  # one launch does both sweeps
        kernel_name = "trx_rollout_kernel<8,7,13> (forward sweep + adjoint sweep, one launch per evaluation)"
    else:
        launch_ms = tape_ms[2] / (args.steps * n_chunks)
        alg_bytes = b_alg / 2  # the adjoint launch must read the per-frame state once
        kernel_name = "trx_vm_staged_kernel (adjoint launch)"
    achieved = alg_bytes * units_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = NCU_DRAM_BYTES_PER_UNIT[ENGINE]
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic * units_per_launch if traffic else None,
        "traffic_source": NCU_SOURCE[ENGINE] + ", per unit x units_per_launch",
        "kernel": kernel_name, "peak_source": peak_source,
        "algorithmic_bytes_per_unit": alg_bytes, "units_per_launch": units_per_launch, "launch_ms": launch_ms,
    }
    if ENGINE == "rollout":
        roofline["checkpoint_bytes_per_unit"] = 2 * obj.info("checkpoint_bytes_per_unit")  # written once, read once
    else:
        roofline["kernel_ms_per_step"] = {"prog": tape_ms[0] / args.steps, "prep_seed": tape_ms[1] / args.steps,
                                          "bprop": tape_ms[2] / args.steps}
    # second roofline (SURVEY.md 8(d) regime ii): fp64 work of the objective, summed over the reference-shaped tapes
    n_ee = 6 if KIND != "chain" else int(os.environ.get("BENCH_CONTACTS", 16))
    n_in, n_out = n_q + 12, n_q + 9 * n_ee
    shapes = [(n_in, HIDDEN), (HIDDEN, n_out)] if HIDDEN > 0 else [(n_in, n_out)]
    flops_unit, uncounted = flops.objective_flops_per_unit(pkg.build_dexsyn, problem, shapes)
    fp64_peak, fp64_source = measured_peak_fp64()
    fp64_achieved = flops_unit * (units / world) / (ms_per_step * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if TOTAL_ROLLOUTS else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_x, "d2h_bytes_per_step": 8 * (n_x + 1)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_fp64": {
            "bound": "fp64", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
            "flops_per_unit": flops_unit, "peak_source": fp64_source, "per": "GPU",
            "flops_source": "sum over the instructions of x_prog + x_prep + x_bprop of one frame (robio_2021_dexopt_b200/flops.py)",
            "uncounted_operators": uncounted,
        },
        "loss": loss,
        "gradient_norm": float(np.linalg.norm(g)),
        "policy_weight_std": WSTD,
        "setup_s": t_build,
        "chunks": n_chunks,
        "device_bytes": obj.info("device_bytes"),
    }
    if parity is not None:
        line["parity_check"] = parity
    # CPU baseline: the reference engine on this box's host cores, bounded sample (N = 1 only)
    if world == 1 and not skip_cpu:
        threads = host_threads()
        r = reference_cpu(threads, CPU_SAMPLE_FRAMES, CPU_SAMPLE_SECONDS, repeats=3)
        if r is not None:
            line["cpu_baseline"] = {
                "value": r["step_rollouts_per_s"], "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": "reference SimpleEngine (oracle/_ref), Batch<double,8>, %d of the %d frames per evaluation, "
                          "%d threads each with its own Memory, median of 3 windows of %.0f s started after all "
                          "threads are ready" % (CPU_SAMPLE_FRAMES, FRAMES, threads, CPU_SAMPLE_SECONDS),
                "windows": r.get("windows")}
            r1 = reference_cpu(1, CPU_SAMPLE_FRAMES, min(CPU_SAMPLE_SECONDS, 5))
            if r1 is not None:
                line["cpu_baseline"]["single_thread_value"] = r1["step_rollouts_per_s"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
