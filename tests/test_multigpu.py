"""Two ranks, two GPUs, rollouts sharded, the all-reduce issued by the native library (trx_comm_* /
trx_objective_set_comm): ten gradient-descent iterates resident on the devices (trx_objective_gd_steps) against the
reference's single-process iterates.  Skipped on a box with one GPU; torch.distributed (gloo) only carries the
communicator id."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import robio_2021_dexopt_b200 as pkg
    from robio_2021_dexopt_b200 import trxfile
    from robio_2021_dexopt_b200.distributed import Communicator
    import rollout_util as ru
    ref = trxfile.load(os.path.join(ROOT, "tests", "golden", name))
    engine = pkg.CudaEngine(rank)
    lanes = 16 // world
    obj = pkg.RolloutObjective(engine, lanes, scalar_rows=(rank == 0), **ru.FIXTURES[name])
    pw, _ = ru.wide_from_tiles(ref, obj.info("n_scalar_rows"))
    obj.parameterize(pw[:, rank * lanes:(rank + 1) * lanes].copy())
    comm = Communicator.from_torch_distributed(rank)
    obj.set_comm(comm)
    # one evaluation through host buffers: every rank returns the job's loss and gradient
    loss, g = obj.evaluate(ref.arrays["x"])
    solver = pkg.GradientDescentSolver(obj, learning_rate=float(ref.arrays["gd/learning_rate"][0]))
    solver.gather(ref.arrays["x"])
    losses = solver.steps_on_device(10)
    results.put((rank, loss, g, losses, solver.scatter()))
    dist.barrier()
    obj.set_comm(None)
    comm.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_gd_iterates_match_reference(golden):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import parity
    name = "rollout_handarm_h64.trxb"
    ref = golden(name)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    want_loss = float(ref.arrays["loss_total"][0])
    for rank, loss, g, losses, x10 in outs:
        assert abs(loss - want_loss) <= 1e-11 * abs(want_loss), (rank, loss, want_loss)
        parity.assert_close(g, ref.arrays["g_total"], "all-reduced gradient on rank %d" % rank)
        want = ref.arrays["gd/loss"]
        assert np.all(np.abs(losses - want) <= 1e-9 * np.abs(want)), (rank, losses, want)
        parity.assert_close(x10, ref.arrays["gd/x"], "x after 10 sharded GD steps, rank %d" % rank, rtol=1e-9, atol=1e-12)
