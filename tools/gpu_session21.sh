#!/bin/bash
# end-of-round evidence refresh (final kernels incl. the short-sequence attention backward): ncu step metrics, bench
# launch list, --set full pages of the fc1-forward GEMM and the short attention backward.  Every ncu command runs AFTER
# the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s21
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
timeout 300 python tools/profile_step.py --warmup 2 --steps 1 > ${O}_plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file ${O}_step_metrics.csv python tools/profile_step.py --warmup 2 --steps 1 > ${O}_ncu_step.log 2>&1; echo "ncu step exit $?"
python tools/summarize_step_metrics.py ${O}_step_metrics.csv > ${O}_step_summary.md 2>&1; head -30 ${O}_step_summary.md
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_bench_plain.json 2>${O}_bench_plain.err; echo "bench plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_ncu_bench.log 2>&1; echo "ncu bench exit $?"
timeout 300 python tools/attn_bench.py > ${O}_attn.log 2>&1; echo "attn bench exit $?"; grep -E "fwd|bwd" ${O}_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_short_kernel -s 30 -c 1 -f -o ${O}_attn_bwd_short python tools/attn_bench.py > ${O}_ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 300 python tools/ncu_one_gemm.py 12608 3072 768 gelu out2 cfg=1 > ${O}_gemm_plain.log 2>&1; echo "gemm plain exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o ${O}_gemm_fc1fwd python tools/ncu_one_gemm.py 12608 3072 768 gelu out2 cfg=1 > ${O}_ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
ls -la gpurun_out | grep s21
