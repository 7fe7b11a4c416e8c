#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s10
timeout 600 python tools/kernel_probe.py gemm > ${O}_probe.log 2>&1; echo "probe exit $?"; grep -c PASS ${O}_probe.log; grep FAIL ${O}_probe.log | head; grep "column" ${O}_probe.log | head -30
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 ${O}_pytest.log | cut -c1-200
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "$name: $(grep -o 'ms_per_step[^,]*' ${O}_bench_${name}.json | head -1)"; }
b raster
b noraster VS_GEMM_RASTER=0
b separate VS_GEMM_COLSUM=separate
b raster2
b separate2 VS_GEMM_COLSUM=separate
timeout 300 python tools/step_breakdown.py > ${O}_breakdown.log 2>&1; grep -E "aux1|GEMM total" ${O}_breakdown.log
