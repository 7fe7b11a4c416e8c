#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s26
BASE=$PWD/visiontransformer_b200/lib/libvitseg_base.so
timeout 400 python tools/kernel_probe.py head loss > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -1 ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
timeout 200 python tools/head_bench.py > ${O}_new.log 2>&1; echo "new exit $?"; cat ${O}_new.log
VS_LIB_PATH=$BASE timeout 200 python tools/head_bench.py > ${O}_base.log 2>&1; echo "base exit $?"; grep -v agreement ${O}_base.log
timeout 600 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_new$i.json 2> ${O}_bench_new$i.err; echo "bench new exit $?"; cut -c1-230 ${O}_bench_new$i.json
VS_LIB_PATH=$BASE timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_base$i.json 2> ${O}_bench_base$i.err; echo "bench base exit $?"; cut -c1-230 ${O}_bench_base$i.json
done
grep -o '"inference": {[^}]*}' ${O}_bench_new1.json ${O}_bench_base1.json
