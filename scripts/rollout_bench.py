"""Quick timing of the re-rolled rollout objective on one GPU (development aid; bench.py is the contract).
usage: python scripts/rollout_bench.py [frames] [lanes] [steps] [key=value problem options ...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robio_2021_dexopt_b200 as pkg  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
opts = dict(kind="handarm", hidden=64)
for a in sys.argv[4:]:
    k, v = a.split("=")
    opts[k] = v
engine = pkg.CudaEngine(0)
t0 = time.time()
max_bytes = int(float(os.environ.get("ROLLOUT_MAX_GB", 0)) * 1e9)
obj = pkg.RolloutObjective(engine, lanes, max_device_bytes=max_bytes, frames=frames, **opts)
print("create %.2fs" % (time.time() - t0), {k: obj.info(k) for k in
      ("n_x", "grid", "rollouts_per_cta", "shared_bytes", "tiles_per_warp", "device_bytes", "n_chunks",
       "checkpoint_bytes_per_unit", "rows_per_frame", "n_joints", "n_levels")})
rng = np.random.default_rng(9)
pw = np.stack([rng.uniform(-0.01, 0.01, lanes), rng.uniform(-0.01, 0.01, lanes), rng.uniform(0, 2 * np.pi, lanes)])
obj.parameterize(pw)
x = pkg.build_dexsyn(frames=1, programs="prog", **opts).arrays["x"]
dev = torch.device("cuda", 0)
d_x = torch.from_numpy(x).to(dev)
d_out = torch.empty(x.size + 1, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
for _ in range(2):
    obj.evaluate_device(d_x.data_ptr(), d_out.data_ptr(), stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    obj.evaluate_device(d_x.data_ptr(), d_out.data_ptr(), stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
out = d_out.cpu().numpy()
print("frames %d lanes %d: %.3f ms per evaluation, %.2f M timestep*rollouts/s, loss %.10g |g| %.6g" %
      (frames, lanes, ms, frames * lanes / ms / 1e3, out[0], np.linalg.norm(out[1:])))

if os.environ.get("ROLLOUT_JSON"):
    import json
    with open(os.environ["ROLLOUT_JSON"], "a") as f:
        f.write(json.dumps({"options": opts, "frames": frames, "rollouts": lanes, "ms_per_evaluation": ms,
                            "timestep_rollouts_per_s": frames * lanes / ms * 1e3, "loss": float(out[0]),
                            "n_chunks": obj.info("n_chunks"), "device_bytes": obj.info("device_bytes"),
                            "checkpoint_bytes_per_unit": obj.info("checkpoint_bytes_per_unit")}) + "\n")
lib = pkg._native.lib()
if hasattr(lib, "trx_rollout_profile_dump"):
    import ctypes
    sys.stdout.flush()
    lib.trx_rollout_profile_dump()
