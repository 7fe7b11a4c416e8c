#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s5
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > ${O}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 ${O}_pytest.log | cut -c1-300
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "$name exit $?: $(cut -c1-200 ${O}_bench_${name}.json)"; }
b streams2 VS_BWD_STREAMS=2
b streams1 VS_BWD_STREAMS=1
b streams2b VS_BWD_STREAMS=2
for cfg in 0 1 2 3 5; do timeout 120 python tools/ncu_one_gemm.py 12608 768 768 f32 res drop cfg=$cfg 2>&1 | tail -1 | cut -c1-300; done
timeout 120 python tools/ncu_one_gemm.py 12608 768 768 f32 res cfg=3 2>&1 | tail -1 | cut -c1-200
timeout 120 python tools/ncu_one_gemm.py 12608 768 768 f32 cfg=3 2>&1 | tail -1 | cut -c1-200
timeout 120 python tools/ncu_one_gemm.py 12608 768 768 cfg=3 2>&1 | tail -1 | cut -c1-200
timeout 120 python tools/ncu_one_gemm.py 12608 768 768 cfg=1 2>&1 | tail -1 | cut -c1-200
echo "== ncu out-proj fwd"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -s 4 -c 1 -o ${O}_gemm_outproj -f python tools/ncu_one_gemm.py 12608 768 768 f32 res drop cfg=3 > ${O}_ncu.log 2>&1; echo "ncu exit $?"; tail -2 ${O}_ncu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -s 4 -c 1 -o ${O}_gemm_outproj_cfg2 -f python tools/ncu_one_gemm.py 12608 768 768 f32 res drop cfg=2 > ${O}_ncu2.log 2>&1; echo "ncu exit $?"
echo "== timeline gaps (1 GPU)"
timeout 600 python tools/dp_timeline.py --out gpurun_out/r02_timeline_n1_c.csv 2>&1 | grep "rank 0"
