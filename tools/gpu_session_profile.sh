#!/bin/bash
# end-of-round evidence: ncu step metrics + bench launch list + sanitizer logs (run AFTER the same commands exited 0 plain)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_prof
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
timeout 300 python tools/profile_step.py --warmup 2 --steps 1 > ${O}_plain.log 2>&1; echo "plain exit $?"
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file ${O}_step_metrics.csv python tools/profile_step.py --warmup 2 --steps 1 > ${O}_ncu_step.log 2>&1; echo "ncu step exit $?"
python tools/summarize_step_metrics.py ${O}_step_metrics.csv > ${O}_step_summary.md 2>&1; head -40 ${O}_step_summary.md
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_bench_plain.json 2>${O}_bench_plain.err; echo "bench plain exit $?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_ncu_bench.log 2>&1; echo "ncu bench exit $?"
echo "== sanitizer"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/kernel_probe.py ln gemm head loss > ${O}_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 ${O}_memcheck.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/kernel_probe.py attn > ${O}_memcheck_attn.log 2>&1; echo "memcheck attn exit $?"; tail -4 ${O}_memcheck_attn.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/kernel_probe.py ln head loss > ${O}_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 ${O}_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/kernel_probe.py ln gemm head loss attn > ${O}_synccheck.log 2>&1; echo "synccheck exit $?"; tail -4 ${O}_synccheck.log
