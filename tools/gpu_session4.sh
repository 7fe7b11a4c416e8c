#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s4
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > ${O}_pytest.log 2>&1
echo "pytest exit $?" | tee -a ${O}_pytest.log
grep -E "passed|failed|FAILED|loss curve|argmax agreement|nvJPEG|class-map" ${O}_pytest.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?: $(cut -c1-200 ${O}_bench.json)"
VS_GEMM_EPI=direct timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_direct.json 2> ${O}_bench_direct.err; echo "bench(direct) exit $?: $(cut -c1-200 ${O}_bench_direct.json)"
timeout 300 python tools/step_breakdown.py > ${O}_breakdown.log 2>&1; tail -42 ${O}_breakdown.log
timeout 300 python tools/gemm_epi_bench.py > ${O}_epi.log 2>&1; grep -E "tile_cfg|fp32 out" ${O}_epi.log
for c in paed_bin paed_multi; do
  timeout 1200 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > ${O}_cfg_${c}.json 2> ${O}_cfg_${c}.err; echo "$c exit $?"; python -c "
import json; d=json.loads(open('${O}_cfg_${c}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['library_baseline'])"
done
