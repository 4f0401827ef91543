"""development aid (GPU): which residual rows of the collision problem differ between the device and the reference
golden, and by how much"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import robio_2021_dexopt_b200 as pkg  # noqa: E402
from robio_2021_dexopt_b200 import trxfile  # noqa: E402

ref = trxfile.load(os.path.join(ROOT, "tests", "golden", "dexsyn_collision.trxb"))
pkg.collision.clear()
mine = pkg.build_dexsyn(frames=3, hidden=4, extra_fixed=4, collision=1, friction=1)
eng = pkg.CudaEngine(0)
prog = mine.view("prog")
params = parity.tile_arrays(ref, "params")
x_prog = eng.compile(mine.programs["prog"])
mem = eng.createMemory(16)
x_prog.parameterize(trxfile.tiles_to_wide(prog.parameters, params), mem)
r = x_prog.run(ref.arrays["x"], mem)
want = trxfile.tiles_to_wide(prog.outputs, parity.tile_arrays(ref, "r"))
err = np.abs(r - want)
bad = np.nonzero(err > 1e-11 * np.maximum(1.0, np.abs(want)))[0]
print("residuals", r.size, "differing", bad.size, "max err", err.max())
layout, total = trxfile.wide_layout(prog.outputs, 16)
for i in bad[:20]:
    port = max(k for k, (off, comps, width) in enumerate(layout) if off <= i)
    off, comps, width = layout[port]
    print("  elem", i, "port", port, "of", len(layout), "comp", (i - off) // width, "lane", (i - off) % width, "got", r[i], "want", want[i])
