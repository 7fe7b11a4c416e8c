#!/bin/bash
# round-2 GPU session 2: PDL, autotuned tiles, pipelined CLC, worker pipeline, headline parity re-run
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s2
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > ${O}_pytest.log 2>&1
echo "pytest exit $?" | tee -a ${O}_pytest.log
grep -E "passed|failed|FAILED|worst pinned|loss curve|argmax agreement|nvJPEG|class-map|p8/1024" ${O}_pytest.log | cut -c1-900
echo "== bench (PDL on, autotune on)"
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err
echo "bench exit $?"; cut -c1-700 ${O}_bench.json; tail -3 ${O}_bench.err
echo "== bench VS_PDL=0"
VS_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_nopdl.json 2> ${O}_bench_nopdl.err
echo "exit $?"; cut -c1-330 ${O}_bench_nopdl.json
echo "== bench VS_GEMM_AUTOTUNE=0"
VS_GEMM_AUTOTUNE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_notune.json 2> ${O}_bench_notune.err
echo "exit $?"; cut -c1-330 ${O}_bench_notune.json
echo "== bench VS_GEMM_SCHED=clc"
VS_GEMM_SCHED=clc timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_clc.json 2> ${O}_bench_clc.err
echo "exit $?"; cut -c1-330 ${O}_bench_clc.json
echo "== gemm microbench static / clc"
timeout 300 python tools/gemm_bench.py > ${O}_gemm_static.log 2>&1
VS_GEMM_SCHED=clc timeout 300 python tools/gemm_bench.py > ${O}_gemm_clc.log 2>&1
paste -d'\n' ${O}_gemm_static.log ${O}_gemm_clc.log | cut -c1-250
echo "== step profile"
timeout 300 python tools/step_breakdown.py > ${O}_breakdown.log 2>&1; tail -45 ${O}_breakdown.log
