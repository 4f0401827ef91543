#!/usr/bin/env python
"""bench.py -- objective+gradient throughput of the tape hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): objective+gradient evaluations per second, counted in timestep*rollouts/s.
One step = one evaluation of the recorded objective and its reverse-mode gradient
(x_prog -> loss -> x_prep -> x_bprop, tractor/include/tractor/solvers/gd.h:50-63) over all T frames of
all R rollouts of the synthetic "Shadow hand + UR arm" problem of BASELINE.json configs[2]:
29 controlled joints, policy MLP 41->64->83, horizon 256, 1024 rollouts per GPU (weak scaling:
configs[3] is 8192 rollouts over 8 GPUs, with one NCCL all-reduce of {loss, gradient} per step).

The tape is recorded by the product's own recorder (csrc/builder) -- nothing under oracle/ is used by the
measured arm.  `--impl reference` and the `cpu_baseline` leg time the reference's own SimpleEngine
(oracle/_ref, built from the reference sources) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES = int(os.environ.get("BENCH_FRAMES", 256))
LANES_PER_GPU = int(os.environ.get("BENCH_LANES", 1024))
KIND = os.environ.get("BENCH_KIND", "handarm")
HIDDEN = int(os.environ.get("BENCH_HIDDEN", 64))
FUSE_DENSE = int(os.environ.get("BENCH_FUSE_DENSE", 1))
FUSED = int(os.environ.get("BENCH_FUSED", 0))
MAX_DEVICE_BYTES = int(float(os.environ.get("BENCH_MAX_DEVICE_GB", 150)) * 1e9)
CPU_SAMPLE_FRAMES = int(os.environ.get("BENCH_CPU_FRAMES", 16))
CPU_SAMPLE_SECONDS = float(os.environ.get("BENCH_CPU_SECONDS", 12))
METRIC = "objective+gradient evals/sec (timestep*rollouts/s)"
UNIT = "timestep*rollouts/s"
# compulsory HBM traffic per timestep*rollout with one state checkpoint written by
# the forward sweep and read back by the adjoint: 2 * 8 * (n_q + 13 * n_dyn) bytes, n_q = 29, n_dyn = 1
B_ALG = 2 * 8 * (29 + 13)
# measured DRAM traffic of the adjoint launch per unit (profiles/README.md): 3.820 GB / (32 frames * 1024 rollouts)
NCU_BPROP_DRAM_BYTES_PER_UNIT = (2.465697e9 + 1.354105e9) / (32 * 1024)


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
            self._stop_flag.wait(0.2)

    def stop(self):
        self._stop_flag.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def reference_cpu(threads, frames, seconds):
    """the reference's own SimpleEngine (oracle/_ref/libtractor_ref.so) on `threads` host threads, each with
    its own Memory over one recording of the same problem; a bounded sample of the workload"""
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
    if not os.path.exists(tool):
        return None
    out = subprocess.run([tool, "bench", "kind=" + KIND, "hidden=%d" % HIDDEN, "frames=%d" % frames,
                          "threads=%d" % threads, "seconds=%g" % seconds], capture_output=True, text=True)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    threads = host_threads()
    t0 = time.time()
    values = []
    total = args.warmup + args.steps
    per = max(2.0, min(20.0, 90.0 / max(1, total)))
    for i in range(total):
        r = reference_cpu(threads, CPU_SAMPLE_FRAMES, per)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_tool is not built"}))
            return 0
        if i >= args.warmup:
            values.append(r["step_rollouts_per_s"])
    value = sum(values) / len(values)
    sample = "reference SimpleEngine, Batch<double,8> lanes, %d frames of the %d-frame rollout per evaluation, %d threads x own Memory, %.0f s per step" % (
        CPU_SAMPLE_FRAMES, FRAMES, threads, per)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * FRAMES * LANES_PER_GPU * args.gpus / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus):
    return {
        "workload": "BASELINE.json configs[2]%s: synthetic Shadow hand + UR arm (29 controlled joints, 60 links, 6 end "
                    "effectors, floating object), policy MLP 41->%d->83, horizon %d, %d rollouts per GPU" % (
                        " sharded as configs[3]" if n_gpus > 1 else "", HIDDEN, FRAMES, LANES_PER_GPU),
        "frames": FRAMES, "rollouts_per_gpu": LANES_PER_GPU, "rollouts": LANES_PER_GPU * n_gpus,
        "parallelism": "rollouts sharded over %d GPU(s); all-reduce of {loss, gradient} per step" % n_gpus if n_gpus > 1
        else "single GPU",
        "dense_layers": "one block instruction per layer (x_dense)" if FUSE_DENSE else "reference expansion (batch + dot4add per weight)",
        "l2": "working set (write-once tape image, tens of GB per chunk) is far larger than the 126 MB L2; no flush needed",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=int(os.environ.get("BENCH_STEPS", 5)))
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

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t_build = time.time()
    bundle = pkg.build_dexsyn(kind=KIND, frames=FRAMES, hidden=HIDDEN, programs="prog,prep,bprop", fuse_dense=FUSE_DENSE)
    engine = pkg.CudaEngine(local_rank)
    obj = pkg.Objective(engine, bundle, lanes=LANES_PER_GPU, max_device_bytes=MAX_DEVICE_BYTES, fused=bool(FUSED),
                        separate=bool(int(os.environ.get("BENCH_SEPARATE", 1))), scalar_rows=(rank == 0))
    t_build = time.time() - t_build
    n_x = obj.n_x
    prog = bundle.view("prog")
    # per-lane random draws: object start offset x,y ~ U(-0.01, 0.01), yaw ~ U(0, 2 pi)  (dexenv_grasp.h:151-164)
    rng = np.random.default_rng(20211227 + 3 + rank)
    n_ports = len(prog.parameters)
    params = np.zeros((n_ports, LANES_PER_GPU))
    params[0] = rng.uniform(-0.01, 0.01, LANES_PER_GPU)
    params[1] = rng.uniform(-0.01, 0.01, LANES_PER_GPU)
    params[2] = rng.uniform(0.0, 2 * np.pi, LANES_PER_GPU)
    obj.parameterize(params.reshape(-1))
    x_host = bundle.arrays["x"].copy()
    units = FRAMES * LANES_PER_GPU * world  # timestep*rollouts per step, whole job

    stream = torch.cuda.current_stream()
    d_x = torch.from_numpy(x_host).cuda()
    d_out = torch.zeros(n_x + 1, dtype=torch.float64, device="cuda")
    h_x = torch.from_numpy(x_host.copy()).pin_memory()
    h_out = torch.zeros(n_x + 1, dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        obj.evaluate_device(d_x.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(d_out)

    def step_e2e():
        if world == 1:
            # the reference-facing C-ABI call with host buffers (trx_objective_evaluate)
            return obj.evaluate(x_host)
        d_x.copy_(h_x, non_blocking=True)
        obj.evaluate_device(d_x.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        dist.all_reduce(d_out)
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return float(h_out[0]), h_out[1:].numpy()

    # ---- device-resident arm: warm-up, then exactly K timed steps
    for _ in range(args.warmup):
        step_device()
    obj.set_profiling(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = pkg.kernel_launch_count()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    kernel_ms = [0.0, 0.0, 0.0]
    for _ in range(args.steps):
        step_device()
        kernel_ms, _n = obj.profile_collect()
    e1.record()
    barrier()
    launches = pkg.kernel_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    obj.set_profiling(False)

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

    ms_per_step = ms_total / args.steps
    value = units / (ms_per_step * 1e-3)
    e2e_value = units / (e2e_s / args.steps)
    peak, peak_source = measured_peak_hbm()
    n_chunks = obj.info("n_chunks")
    bprop_ms_per_launch = kernel_ms[2] / (args.steps * n_chunks)
    units_per_launch = FRAMES * obj.info("chunk_lanes")
    # the adjoint kernel's share of the algorithmic bytes: it must read the per-frame state once (B_ALG / 2)
    achieved = (B_ALG / 2) * units_per_launch / (bprop_ms_per_launch * 1e-3) / 1e9
    tape_bytes = obj.info("prog.arg_bytes_per_tile") + obj.info("prep.arg_bytes_per_tile") + obj.info(
        "bprop.arg_bytes_per_tile")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_x, "d2h_bytes_per_step": 8 * (n_x + 1)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of the adjoint launch from the committed ncu --set full
            # capture (profiles/ncu_full_r1_staged_T32_raw.csv: 3.820 GB for 32 frames x 1024 rollouts), scaled
            # to this launch's units; only for the default dense-fused generic lowering it was captured on
            "traffic": (NCU_BPROP_DRAM_BYTES_PER_UNIT * units_per_launch if (not FUSED and FUSE_DENSE) else None),
            "traffic_source": "profiles/ncu_full_r1_staged_T32_raw.csv, per unit x units_per_launch",
            "kernel": ("trx_vm_kernel" if FUSED else "trx_vm_staged_kernel") + " (adjoint launch)", "lowering": "fused (chain schedule + slot rings)" if FUSED else "generic (levels of chains; stream pages staged in shared memory)", "peak_source": peak_source,
            "algorithmic_bytes_per_unit": B_ALG / 2, "units_per_launch": units_per_launch,
            "launch_ms": bprop_ms_per_launch,
            "kernel_ms_per_step": {"prog": kernel_ms[0] / args.steps, "prep_seed": kernel_ms[1] / args.steps,
                                   "bprop": kernel_ms[2] / args.steps},
            "interpreted_tape_bytes_per_unit": tape_bytes / (8.0 * FRAMES),
            "interpreted_tape_gbs": tape_bytes / (8.0 * FRAMES) * units / world / (sum(kernel_ms) / args.steps * 1e-3) / 1e9,
        },
        "loss": loss,
        "lowering_stats": {k: {q: obj.info("%s.%s" % (k, q)) for q in ("n_instructions", "n_levels", "n_bundles", "stream_words", "n_ring_reads", "n_image_reads", "n_ring_only_writes", "n_dual_writes", "n_image_writes", "n_prefetch_lines", "n_chained", "n_folded")} for k in ("prog", "prep", "bprop")},
        "direct_operands": {k: obj.info("direct." + k) for k in ("n_prepare_dropped", "n_arguments_redirected", "n_mul_rewritten")},
        "setup_s": t_build,
        "chunks": n_chunks,
        "device_bytes": obj.info("device_bytes"),
    }
    # CPU baseline: the reference engine on this box's host cores, bounded sample (N = 1 only)
    if world == 1 and os.environ.get("BENCH_SKIP_CPU", "0") != "1":
        threads = host_threads()
        r = reference_cpu(threads, CPU_SAMPLE_FRAMES, CPU_SAMPLE_SECONDS)
        if r is not None:
            line["cpu_baseline"] = {
                "value": r["step_rollouts_per_s"], "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": "reference SimpleEngine (oracle/_ref), Batch<double,8>, %d of the %d frames per evaluation, "
                          "%d threads each with its own Memory, %.0f s" % (CPU_SAMPLE_FRAMES, FRAMES, threads,
                                                                           CPU_SAMPLE_SECONDS)}
            r1 = reference_cpu(1, CPU_SAMPLE_FRAMES, min(CPU_SAMPLE_SECONDS, 6))
            if r1 is not None:
                line["cpu_baseline"]["single_thread_value"] = r1["step_rollouts_per_s"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
