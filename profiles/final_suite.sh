#!/bin/bash
# Round-2 measurement suite, run on the GPU box in one call: the bench lines, the reference arm, the ncu launch list and the
# ncu --set full captures the profiles/ summaries are made from.  Outputs under gpurun_out/.
set -x
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference.json 2> /dev/null
timeout 400 python bench.py --rotation-ratio none --no-cpu-baseline --no-e2e --no-configs --no-cfg5 > $O/r2_bench_general.json 2> /dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cudnn --no-variants --no-configs --no-cfg5 --no-floor > $O/ncu_launch.log 2>&1
for k in stn_fwd_kernel stn_bwd_band; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 12 -c 1 -f -o $O/r2_${k}_cfg2 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cudnn --no-variants --no-configs --no-cfg5 --no-floor > $O/ncu_$k.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:stn_bwd_band --launch-skip 2 -c 1 -f -o $O/r2_stn_bwd_band_cfg5 python profiles/run_bwd_once.py cfg5 > $O/ncu_band_cfg5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:kframe --launch-skip 2 -c 1 -f -o $O/r2_stn_bwd_kframe_cfg3 python profiles/run_bwd_once.py cfg3 > $O/ncu_kframe_cfg3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:kframe --launch-skip 2 -c 1 -f -o $O/r2_stn_bwd_kframe_cfg4 python profiles/run_bwd_once.py cfg4 > $O/ncu_kframe_cfg4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:theta_tab --launch-skip 2 -c 1 -f -o $O/r2_stn_bwd_theta_tab_cfg4 python profiles/run_bwd_once.py cfg4 > $O/ncu_theta_cfg4.log 2>&1
timeout 200 python profiles/kframe_time.py cfg4 > $O/r2_kframe_time_cfg4.jsonl 2>&1
KF_ROWS=0,16,32,64,112 timeout 200 python profiles/kframe_time.py cfg3 > $O/r2_kframe_time_cfg3.jsonl 2>&1
timeout 100 python profiles/ingest_time.py > $O/r2_ingest_time.jsonl 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:ingest --launch-skip 2 -c 2 -f -o $O/r2_ingest python profiles/run_ingest_once.py > $O/ncu_ingest.log 2>&1
ls -la $O | tail -30
