#!/bin/bash
# Round-end measurement suite, run on the GPU box: bench lines for every BASELINE config, the reference arm, the ncu
# launch list and the ncu --set full captures the profiles/ summaries are made from.
set -x
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/bench_r1.json 2> $O/bench_r1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r1_reference.json 2> $O/bench_r1_reference.err
for c in cfg1 cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $c --no-cpu-baseline --no-e2e > $O/bench_r1_$c.json 2> $O/bench_r1_$c.err
done
timeout 400 python bench.py --workload cfg3 --band off --no-cpu-baseline --no-e2e --no-cudnn > $O/bench_r1_cfg3_general.json 2> /dev/null
timeout 400 python bench.py --no-gx --no-cpu-baseline --no-e2e --no-cudnn > $O/bench_r1_nogx.json 2> /dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cudnn --no-variants > $O/ncu_launch.log 2>&1
for k in stn_fwd_kernel stn_bwd_kernel prepare_images_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 12 -c 1 -f -o $O/r1_$k \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cudnn > $O/ncu_$k.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:stn_bwd_band --launch-skip 2 -c 1 -f -o $O/r1_stn_bwd_band_cfg3 \
  python profiles/run_bwd_once.py cfg3 > $O/ncu_band_cfg3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:stn_bwd_band --launch-skip 2 -c 1 -f -o $O/r1_stn_bwd_rowband_cfg2 \
  python profiles/run_bwd_once.py cfg2 > $O/ncu_rowband_cfg2.log 2>&1
ls -la $O | tail -20
