#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s3
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > ${O}_pytest.log 2>&1
echo "pytest exit $?" | tee -a ${O}_pytest.log
grep -E "passed|failed|FAILED|loss curve|argmax agreement|nvJPEG|class-map|p8/1024/16h P" ${O}_pytest.log | cut -c1-600
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "$name exit $?: $(cut -c1-200 ${O}_bench_${name}.json)"; }
b static VS_GEMM_SCHED=static
b clc VS_GEMM_SCHED=clc
b static_pdl VS_GEMM_SCHED=static VS_PDL=1
b static2 VS_GEMM_SCHED=static
echo "== gemm microbench static / clc (no PDL)"
timeout 300 python tools/gemm_bench.py > ${O}_gemm_static.log 2>&1
VS_GEMM_SCHED=clc timeout 300 python tools/gemm_bench.py > ${O}_gemm_clc.log 2>&1
paste -d'\n' ${O}_gemm_static.log ${O}_gemm_clc.log | cut -c1-250
echo "== step profile"
timeout 300 python tools/step_breakdown.py > ${O}_breakdown.log 2>&1; tail -42 ${O}_breakdown.log
echo "== full bench (cpu + library baselines)"
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_full.json 2> ${O}_bench_full.err; echo "exit $?"; cat ${O}_bench_full.json | cut -c1-3000
for c in paed_bin paed_multi vitl384 infer512; do
  echo "== bench --config $c"
  timeout 1200 python bench.py --config $c --steps 10 --warmup 3 > ${O}_cfg_${c}.json 2> ${O}_cfg_${c}.err; echo "exit $?"; cut -c1-1500 ${O}_cfg_${c}.json; tail -2 ${O}_cfg_${c}.err
done
