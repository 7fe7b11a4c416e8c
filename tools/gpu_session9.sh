#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s9
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "$name: $(cut -c1-150 ${O}_bench_${name}.json | grep -o 'ms_per_step.*' )"; }
b fused
b separate VS_GEMM_COLSUM=separate
b fused2
b separate2 VS_GEMM_COLSUM=separate
timeout 300 python tools/step_breakdown.py > ${O}_breakdown_fused.log 2>&1; grep -E "aux1|colsum|GEMM total" ${O}_breakdown_fused.log
VS_GEMM_COLSUM=separate timeout 300 python tools/step_breakdown.py > ${O}_breakdown_sep.log 2>&1; grep -E "aux1|colsum|GEMM total" ${O}_breakdown_sep.log
