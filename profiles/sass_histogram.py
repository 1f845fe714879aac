"""Opcode histogram per kernel of the built library (cuobjdump -sass; no GPU needed): evidence in the tree of what the
kernels are made of -- TMA bulk copies (UBLKCP), tensor stores (UTMASTG), cluster barriers / DSMEM (UCGABAR, mapa -> ...),
programmatic dependent launch (ACQBULK / griddepcontrol -> ...), shared-memory traffic, no tensor-core instructions.
usage: python profiles/sass_histogram.py [regex of demangled kernel names] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "loans_b200", "libloans_stn.so")
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else ".")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, hist = None, {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(stn::CropParams.*\)$", "(...)", name).replace("stn::", "")
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", ln)
    if m and name:
        op = m.group(1)
        base = op.split(".")[0]
        key = op if base in ("UBLKCP", "UTMASTG", "UTMALDG", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "ATOMS", "RED", "REDG", "ATOMG", "BAR",
                             "ACQBULK", "CCTL", "MEMBAR", "FENCE", "ERRBAR", "LDGSTS", "UTCHMMA", "UTCQMMA", "HMMA", "IMMA") else base
        hist[name][key] += 1
print("# SASS opcode histogram per kernel of loans_b200/libloans_stn.so (sm_100a), `cuobjdump -sass`; instructions by static count")
for k in sorted(hist):
    if not want.search(k) or not hist[k]:
        continue
    h = hist[k]
    tot = sum(h.values())
    notable = {o: c for o, c in h.items() if re.match(r"UBLKCP|UTMA|SYNCS|UCGABAR|ACQBULK|ATOMS|RED|ATOMG|LDGSTS|UTC|HMMA|IMMA|FENCE|MEMBAR|BAR", o)}
    top = ", ".join("%s %d" % (o, c) for o, c in h.most_common(12))
    print("\n%s\n  %d instructions; %s" % (k, tot, top))
    if notable:
        print("  async / sync / atomics: " + ", ".join("%s %d" % (o, c) for o, c in sorted(notable.items())))
    tens = [o for o in h if re.match(r"UTC|HMMA|IMMA|QMMA|DMMA", o)]
    print("  tensor-core instructions: %s" % (", ".join(tens) if tens else "none"))
