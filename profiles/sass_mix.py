"""Summarise an `ncu --page source --csv` dump: instruction mix and execution-count profile of the first kernel."""
import collections
import csv
import itertools
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[0] != "Address":
        data.append(r)
ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ist = hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data)
print("kernel:", rows[0][1])
print("total warp instructions", tot, " SASS lines", len(data))
by = collections.Counter(); bys = collections.Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    by[op] += int(r[ia]); bys[op] += int(r[ist])
for op, c in by.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print("%-10s %10d %5.1f%%  stall samples %d" % (op, c, 100.0 * c / tot, bys[op]))
cnts = [int(r[ia]) for r in data]
seg = []
for k, g in itertools.groupby(enumerate(cnts), key=lambda t: t[1]):
    g = list(g); seg.append((g[0][0], g[-1][0], k))
print("execution-count segments (first, last SASS line, warp executions):")
print(seg[:60])
