"""Joins the per-instruction stall samples of an ncu capture (ncu -i X.ncu-rep --page source --csv) with the line
table of the same build (nvdisasm -g -c of the kernel's cubin) and prints the samples by stall reason and by
innermost source line.  Development aid; profiles/ncu_r2_rollout_stalls_by_line.txt was written by it.
usage: python scripts/devtools/ncu_source_lines.py source.csv listing.sass <mangled kernel name substring>"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
lines = open(sys.argv[2]).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and sys.argv[3] in l][0]
cur, off2line = ("?", 0), {}
for l in lines[start + 1:]:
    if (l.startswith("\t.section") or l.startswith("//-----")) and off2line:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
assert len(off2line) == len(data), "listing and capture are different builds"
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
S, IE = col["# Samples"], col["Instructions Executed"]
by_line, inst, reason, line_reason = collections.Counter(), collections.Counter(), collections.Counter(), {}
for r in data:
    ln = off2line[int(r[0], 16) - base]
    by_line[ln] += int(r[S] or 0)
    inst[ln] += int(r[IE] or 0)
    for h in stalls:
        v = int(r[col[h]] or 0)
        reason[h] += v
        line_reason.setdefault(ln, collections.Counter())[h] += v
tot = sum(by_line.values())
print("warp stall samples: %d; warp instructions executed: %.3f G" % (tot, sum(inst.values()) / 1e9))
print("by reason: " + ", ".join("%s %.1f %%" % (h[6:], 100 * v / tot) for h, v in reason.most_common(10)))
nb = tot - reason["stall_barrier"]
print("without the barrier samples (%d): " % nb + ", ".join(
    "%s %.1f %%" % (h[6:], 100 * v / nb) for h, v in reason.most_common(10) if h != "stall_barrier"))
print()
print("%-8s %-26s %12s  %s" % ("samples", "innermost source line", "warp-inst", "top reasons"))
for ln, n in by_line.most_common(40):
    top = line_reason[ln].most_common(3)
    print("%6.2f %% %-26s %12d  %s" % (100 * n / tot, "%s:%d" % ln, inst[ln], " ".join("%s %d" % (h[6:], v) for h, v in top)))
