#!/bin/bash
# A/B helper run on the GPU box: rebuild libloans_stn.so with the given EXTRA flags and print the bench's key numbers.
# usage: profiles/ab.sh "<label>" "<EXTRA nvcc flags>" [bench args...]
label="$1"; extra="$2"; shift 2
make -C loans_b200/csrc clean >/dev/null 2>&1
make -C loans_b200/csrc -j8 EXTRA="$extra" >/dev/null 2>&1 || { echo "$label: build failed"; exit 1; }
timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err || { echo "$label: bench failed"; tail -3 gpurun_out/ab_tmp.err; exit 1; }
python - "$label" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_tmp.json"))
r = d["roofline"]
print("%-28s step %.2f us  fwd %.2f us  bwd %.2f us  value %.3g crops/s  step-frac %.3f  bwd-frac %.3f" % (
    sys.argv[1], d["ms_per_step"] * 1e3, r["fwd_kernel"]["avg_launch_us"], r["avg_launch_us"], d["value"],
    r["whole_step"]["frac"], r["frac"]))
PY
