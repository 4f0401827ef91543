"""Per-barrier-interval view of an `ncu --page source --csv` export of the rollout kernel (development aid).

usage: python scripts/devtools/ncu_intervals.py <source.csv> <trx_rollout.o> <kernel mangled-name fragment>

Cuts the kernel's SASS at every BAR.SYNC and reports, per interval, the sampled warp states (working samples by
stall reason, samples of warps waiting at the closing barrier), executed warp instructions and the source
functions the interval's instructions come from (nvdisasm line info of the same object file)."""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

src_csv, obj, frag = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
sec = [i for i, l in enumerate(sass) if l.startswith("//---") and ".text." in l]
tgt = [i for i in sec if frag in sass[i]][0]
end = ([i for i in sec if i > tgt] + [len(sass)])[0]
ins, cur = [], None
for l in sass[tgt:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins.append(cur)


def functions(path):
    out = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"(?:template <[^>]*>\s*)?(?:RL_HD|TRX_HD|__device__ __forceinline__|__global__) [\w:]+ \*?(\w+)\(", l)
        if m:
            out.append((i, m.group(1)))
    return out


root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "robio_2021_dexopt_b200", "csrc")
F = {"rollout_frame.h": functions(os.path.join(root, "rollout", "rollout_frame.h"))}


def where(c):
    if c is None:
        return "?"
    f, l = c
    if f in F:
        prev = [n for (s, n) in F[f] if s <= l]
        return prev[-1] if prev else f
    if f == "trx_rollout.cu":
        return "cu:%d" % (l // 10 * 10)
    return f.split(".")[0]


rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(ins), (len(data), len(ins))
st = ["stall_no_inst", "stall_wait", "stall_short_sb", "stall_selected", "stall_branch_resolving", "stall_math",
      "stall_long_sb", "stall_barrier", "stall_not_selected", "stall_dispatch", "stall_mio"]
segs, cur = [], None


def fresh(first):
    d = {"n": 0, "first": first, "ex": 0, "thr": 0, "fn": collections.Counter()}
    d.update({s: 0 for s in st})
    return d


cur = fresh(0)
for i, r in enumerate(data):
    for s in st:
        cur[s] += int(r[ci[s]])
    cur["ex"] += int(r[ci["Instructions Executed"]])
    cur["thr"] += int(r[ci["Thread Instructions Executed"]])
    cur["n"] += 1
    cur["fn"][where(ins[i])] += int(r[ci["# Samples"]]) - int(r[ci["stall_barrier"]])
    if "BAR.SYNC" in r[ci["Source"]]:
        cur["last"] = i
        segs.append(cur)
        cur = fresh(i + 1)
segs.append(cur)
tot = sum(sum(s[k] for k in st) for s in segs)
print("total samples", tot)
for s in segs:
    work = sum(s[k] for k in st if k != "stall_barrier")
    if work + s["stall_barrier"] < tot / 400:
        continue
    top = ", ".join("%s" % k for k, _ in s["fn"].most_common(3))
    print("instr %5d-%5d n=%4d exec %7dk lanes %4.1f | wall %4.1f%% | work %6d: noi %3.0f%% wait %3.0f%% ssb %3.0f%% sel %3.0f%% | barrier %6d | %s" % (
        s["first"], s.get("last", 0), s["n"], s["ex"] // 1000, s["thr"] / max(1, s["ex"]),
        100 * (work + s["stall_barrier"]) / tot, work, 100 * s["stall_no_inst"] / max(1, work),
        100 * s["stall_wait"] / max(1, work), 100 * s["stall_short_sb"] / max(1, work),
        100 * s["stall_selected"] / max(1, work), s["stall_barrier"], top))
