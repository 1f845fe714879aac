"""Attribute an ncu SASS-level source dump (ncu --page source --csv) to CUDA source lines using nvdisasm -g output
of the same kernel (instruction order is identical).  usage: line_profile.py <ncu_sass.csv> <nvdisasm.sass> [min_pct]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[0] != "Address":
        data.append(r)
ia = hdr.index("Instructions Executed"); ist = hdr.index("# Samples")
lines = []          # (file, line) per instruction, in order
cur = ("?", 0)
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
print("ncu instructions", len(data), " nvdisasm instructions", len(lines))
n = min(len(data), len(lines))
inst = collections.Counter(); samp = collections.Counter()
for k in range(n):
    inst[lines[k]] += int(data[k][ia]); samp[lines[k]] += int(data[k][ist])
ti, ts = sum(inst.values()), sum(samp.values())
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
print("total warp instr %d, stall samples %d" % (ti, ts))
for key in sorted(inst, key=lambda k: (k[0], k[1])):
    pi, ps = 100.0 * inst[key] / ti, 100.0 * samp[key] / max(ts, 1)
    if pi >= minp or ps >= minp:
        print("%-18s:%4d  instr %5.1f%%  stalls %5.1f%%" % (key[0], key[1], pi, ps))
