"""GPU probe (development aid): time the generic tape VM on a reference-recorded dexsyn tape.
Uses oracle/_ref/ref_tool to record the tape on the box (test infrastructure), so it is NOT bench.py."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import robio_2021_dexopt_b200 as pkg
from robio_2021_dexopt_b200 import trxfile

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
path = "/tmp/dex%d.trxb" % frames
t0 = time.time()
subprocess.check_call([os.path.join(ROOT, "oracle/_ref/ref_tool"), "record", path, "problem=dexsyn",
                       "frames=%d" % frames, "tiles=1"], stderr=subprocess.DEVNULL)
print("recorded in %.1fs" % (time.time() - t0))
b = trxfile.load(path)
eng = pkg.CudaEngine(0)
t0 = time.time()
xs = {n: eng.compile(b.programs[n]) for n in ("prog", "prep", "bprop")}
print("compiled in %.1fs" % (time.time() - t0))
for n, x in xs.items():
    print(n, {k: x.stat(k) for k in ("n_instructions", "n_levels", "n_bundles", "stream_words", "memory_size", "max_level_width", "arg_bytes_per_tile")})
prog = b.view("prog")
tiles = lanes // 8
mem = eng.createMemory(lanes)
params = trxfile.tiles_to_wide(prog.parameters, [b.arrays["tile0/params"]] * tiles)
xs["prog"].parameterize(params, mem)
x = b.arrays["x"]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(3):
    e0 = ev(); xs["prog"].input(x, mem); xs["prog"].execute(mem); e1 = ev()
    r = xs["prog"].output(mem); e2 = ev()
    xs["prep"].execute(mem); e3 = ev()
    xs["bprop"].input(r, mem); xs["bprop"].execute(mem); e4 = ev()
    g = xs["bprop"].output(mem); e5 = ev()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(bb) for a, bb in ((e0, e1), (e1, e2), (e2, e3), (e3, e4), (e4, e5))]
    tot = e0.elapsed_time(e5)
    print("iter %d: prog %.2f out %.2f prep %.2f bprop %.2f gout %.2f total %.2f ms -> %.3g step*rollouts/s; mem %.1f GB" %
          (it, *ms, tot, frames * lanes / (tot * 1e-3), mem.bytes() / 1e9))
per = trxfile.wide_to_tiles(prog.outputs, r, tiles)
import parity
parity.assert_close(per[0], b.arrays["tile0/r"], "r tile0")
parity.assert_close(per[-1], b.arrays["tile0/r"], "r last tile")
print("parity ok; loss", float(r @ r), "g norm", float(np.linalg.norm(g)))
