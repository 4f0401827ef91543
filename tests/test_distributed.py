"""world_size-2 gloo test of the multi-GPU path's host logic: rollouts sharded by lane tiles, the
lane-independent residual rows counted by rank 0 only, one all-reduce(sum) of {loss, gradient}."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conftest
    import parity
    from emu_engine import EmuEngine
    from robio_2021_dexopt_b200 import trxfile
    from robio_2021_dexopt_b200.distributed import allreduce_loss_gradient, shard_lanes, shard_wide_parameters
    bundle = trxfile.load(os.path.join(ROOT, "tests", "golden", "dexsyn_small.trxb.xz"))
    prog = bundle.view("prog")
    engine = EmuEngine(conftest._build_emulator())
    total = 16
    first, count = shard_lanes(total, rank, world)
    wide_all = trxfile.tiles_to_wide(prog.parameters, parity.tile_arrays(bundle, "params"))
    x_prog = engine.compile(bundle.programs["prog"])
    x_prep = engine.compile(bundle.programs["prep"])
    x_bprop = engine.compile(bundle.programs["bprop"])
    mem = engine.createMemory(count)
    x_prog.parameterize(shard_wide_parameters(prog.parameters, wide_all, total, first, count), mem)
    r = x_prog.run(bundle.arrays["x"], mem)
    if rank != 0:  # what TRX_OBJECTIVE_NO_SCALAR_ROWS does on the device
        off = 0
        for p in prog.outputs:
            n = p.size // 8 if p.lanes == 1 else (p.size // 64) * count
            if p.lanes == 1:
                r[off:off + n] = 0.0
            off += n
    x_prep.execute(mem)
    g = x_bprop.run(r, mem)
    buf = torch.from_numpy(np.concatenate([[float(r @ r)], g]))
    allreduce_loss_gradient(buf)
    if rank == 0:
        results.put(buf.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_gradient(golden):
    import parity
    from robio_2021_dexopt_b200.distributed import shard_lanes
    assert shard_lanes(8192, 3, 8) == (3072, 1024)
    assert [shard_lanes(40, r, 3)[1] for r in range(3)] == [16, 16, 8]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    bundle = golden("dexsyn_small.trxb.xz")
    want_loss = float(bundle.arrays["loss_total"][0])
    assert abs(out[0] - want_loss) <= 1e-11 * abs(want_loss)
    parity.assert_close(out[1:], bundle.arrays["g_total"], "all-reduced gradient")
