#!/bin/bash
# EPI 1 chunk loop rolled (instruction-cache footprint) vs the unrolled build (VS_LIB_PATH = libvitseg_base.so)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s22
BASE=$PWD/visiontransformer_b200/lib/libvitseg_base.so
timeout 400 python tools/kernel_probe.py gemm > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -2 ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
EPI_CFGS=1 timeout 200 python tools/gemm_epi_bench.py > ${O}_epi_new.log 2>&1; echo "epi new exit $?"; cat ${O}_epi_new.log
VS_LIB_PATH=$BASE EPI_CFGS=1 timeout 200 python tools/gemm_epi_bench.py > ${O}_epi_base.log 2>&1; echo "epi base exit $?"; cat ${O}_epi_base.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_new$i.json 2> ${O}_bench_new$i.err; echo "bench new exit $?"; cut -c1-260 ${O}_bench_new$i.json
VS_LIB_PATH=$BASE timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_base$i.json 2> ${O}_bench_base$i.err; echo "bench base exit $?"; cut -c1-260 ${O}_bench_base$i.json
done
grep -o '"inference": {[^}]*}' ${O}_bench_new1.json
