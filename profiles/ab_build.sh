#!/bin/bash
# A/B helper run on the GPU box: for each set of EXTRA nvcc flags, rebuild the library and time the backward kernels.
# usage: profiles/ab_build.sh "cfg2 cfg5" "<flags 1>" "<flags 2>" ...
wls="$1"; shift
for extra in "$@"; do
  make -C loans_b200/csrc clean >/dev/null 2>&1
  make -C loans_b200/csrc -j8 EXTRA="$extra" >/dev/null 2>&1 || { echo "[$extra] build failed"; continue; }
  echo "== EXTRA=[$extra]"
  SWEEP="${AB_SWEEP:-}" timeout 300 python profiles/band_sweep.py $wls 2>&1 | tail -${AB_TAIL:-8}
done
