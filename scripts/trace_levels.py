"""development aid: per-level timing of the generic tape VM on tile 0 (clock64 at each barrier)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import robio_2021_dexopt_b200 as pkg
from robio_2021_dexopt_b200 import _native, trxfile
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
which = sys.argv[3] if len(sys.argv) > 3 else "prog"
b = pkg.build_dexsyn(frames=frames, fuse_dense=1, programs="prog,prep,bprop")
eng = pkg.CudaEngine(0)
xs = {n: eng.compile(b.programs[n]) for n in ("prog", "prep", "bprop")}
mem = eng.createMemory(lanes)
prog = b.view("prog")
rng = np.random.default_rng(0)
params = np.zeros((len(prog.parameters), lanes)); params[2] = rng.uniform(0, 6.28, lanes)
xs["prog"].parameterize(params.reshape(-1), mem)
L = _native.lib()
L.trx_debug_trace.argtypes = [ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]
def run(name):
    if name == "prog": return xs["prog"].run(b.arrays["x"], mem)
    if name == "prep": xs["prep"].execute(mem); return None
for it in range(2):
    r = xs["prog"].run(b.arrays["x"], mem); xs["prep"].execute(mem); g = xs["bprop"].run(r, mem)
nlev = xs[which].stat("n_levels"); W = xs[which].stat("n_warps")
ptr = ctypes.c_void_p()
NW = (1 << 20) + 32 * 4096 * 4 + 64
assert nlev * W * 2 < (1 << 20) - 1
L.trx_debug_trace(NW, ctypes.byref(ptr))
flag = torch.ones(1, dtype=torch.int64, device='cuda')
ctypes.CDLL('libcudart.so').cudaMemcpy(ctypes.c_void_p(ptr.value + ((1 << 20) - 1) * 8), ctypes.c_void_p(flag.data_ptr()), ctypes.c_size_t(8), 3)
if which == "prog": xs["prog"].run(b.arrays["x"], mem)
elif which == "bprop": xs["bprop"].run(r, mem)
torch.cuda.synchronize()
buf = torch.empty(nlev * W * 2, dtype=torch.int64, device="cuda")
ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ptr, ctypes.c_size_t(nlev * W * 16), 3)
t = buf.cpu().numpy().reshape(nlev, W, 2)
recbuf = torch.empty(32 * 4096 * 4, dtype=torch.int64, device="cuda")
ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(recbuf.data_ptr()), ctypes.c_void_p(ptr.value + (1 << 20) * 8), ctypes.c_size_t(32 * 4096 * 32), 3)
rec = recbuf.cpu().numpy().reshape(32, 4096, 4)
dclk = torch.empty(8, dtype=torch.int64, device='cuda')
ctypes.CDLL('libcudart.so').cudaMemcpy(ctypes.c_void_p(dclk.data_ptr()), ctypes.c_void_p(ptr.value + ((1 << 20) + 32 * 4096 * 4) * 8), ctypes.c_size_t(64), 3)
dclk = dclk.cpu().numpy()
print('raw dense clocks', dclk.tolist())
if dclk[7]: print('dense reverse, mean clocks per phase {preload issue, stage+barrier, gx, gW, gb+barrier}:', (dclk[:5] / dclk[7]).round(0).tolist(), 'calls', int(dclk[7]))
L.trx_debug_trace(0, None)
valid = rec[:, :, 0] > 0
args_t = (rec[:, :, 1] - rec[:, :, 0])[valid]
exec_t = (rec[:, :, 2] - rec[:, :, 1])[valid]
ops = (rec[:, :, 3] & 0xffff)[valid]
print("records traced", valid.sum(), "arg fetch cycles mean/median/p90", args_t.mean(), np.median(args_t), np.percentile(args_t, 90), "exec mean/median/p90", exec_t.mean(), np.median(exec_t), np.percentile(exec_t, 90))
names = [L.trx_op_name(i).decode() for i in range(L.trx_op_count())]
import collections
by = collections.defaultdict(list)
for o, a, e in zip(ops, args_t, exec_t): by[int(o)].append((a, e))
for o, v in sorted(by.items(), key=lambda kv: -sum(a + e for a, e in kv[1]))[:14]:
    v = np.array(v); print("  %-34s n=%5d args %6.0f exec %7.0f" % (names[o] if o < len(names) else o, len(v), v[:, 0].mean(), v[:, 1].mean()))
arrive, leave = t[:, :, 0], t[:, :, 1]
rel = leave.max(axis=1)
dur = np.diff(rel)
busy = arrive[1:] - leave[:-1]            # per warp: time from leaving previous barrier to arriving at this one
wait = leave - arrive
print("per-warp mean busy", busy.mean(axis=0).round(0).tolist())
print("per-warp mean wait", wait.mean(axis=0).round(0).tolist())
print("levels", nlev, "total cycles", rel[-1] - rel[0], "mean level", dur.mean(), "median", np.median(dur), "p90", np.percentile(dur, 90), "max", dur.max())
print("mean slowest-warp busy", busy.max(axis=1).mean(), "mean avg-warp busy", busy.mean(), "mean barrier overhead (level dur - slowest busy)", (dur - busy.max(axis=1)).mean())
idle_frac = 1 - busy.clip(min=0).sum() / (dur.sum() * W)
print("fraction of warp-time waiting at barriers", idle_frac)
hist, edges = np.histogram(dur, bins=[0, 500, 1000, 2000, 3000, 5000, 10000, 20000, 50000, 1e9])
for h, e0, e1 in zip(hist, edges[:-1], edges[1:]): print("  %7d..%7d cycles: %6d levels, %5.1f%% of time" % (e0, min(e1, 9999999), h, 100 * dur[(dur >= e0) & (dur < e1)].sum() / dur.sum()))
info = np.zeros(nlev, dtype=np.uint32)
L.trx_debug_level_info.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
L.trx_debug_level_info(xs[which]._h, info.ctypes.data, nlev)
nb, nc = (info & 0xffff)[1:], (info >> 16)[1:]
print("levels with coop records: %d, share of time %.1f%%" % ((nc > 0).sum(), 100 * dur[nc > 0].sum() / dur.sum()))
for lo, hi in ((0, 1), (1, 2), (2, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 100000)):
    sel = (nb >= lo) & (nb < hi) & (nc == 0)
    if sel.any(): print("  %5d..%5d bundles: %5d levels, mean %7.0f cycles, %5.1f%% of time" % (lo, hi, sel.sum(), dur[sel].mean(), 100 * dur[sel].sum() / dur.sum()))
sel = nc > 0
print("  coop levels: mean %7.0f cycles; bundles alongside mean %.1f" % (dur[sel].mean(), nb[sel].mean()))
np.save(os.path.join(ROOT, "gpurun_out", "trace_%s.npy" % which), t)
