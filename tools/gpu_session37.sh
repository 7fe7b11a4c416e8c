#!/bin/bash
# final evidence capture (end-of-round build): ncu step metrics, bench launch list, --set full page of the persistent short attention forward
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s37
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
timeout 300 python tools/profile_step.py --warmup 2 --steps 1 > ${O}_plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file ${O}_step_metrics.csv python tools/profile_step.py --warmup 2 --steps 1 > ${O}_ncu_step.log 2>&1; echo "ncu step exit $?"
python tools/summarize_step_metrics.py ${O}_step_metrics.csv > ${O}_step_summary.md 2>&1; head -30 ${O}_step_summary.md
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_bench_plain.json 2>${O}_bench_plain.err; echo "bench plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_ncu_bench.log 2>&1; echo "ncu bench exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_short_persist_kernel -s 30 -c 1 -f -o ${O}_attn_fwd_persist python tools/attn_bench.py > ${O}_ncu_attn.log 2>&1; echo "ncu attn exit $?"
