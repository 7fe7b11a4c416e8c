#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s25
BASE=$PWD/visiontransformer_b200/lib/libvitseg_base.so
for i in 1 2; do
timeout 200 python tools/head_bench.py > ${O}_new$i.log 2>&1; echo "new $i exit $?"; grep "cold L2" ${O}_new$i.log
VS_LIB_PATH=$BASE timeout 200 python tools/head_bench.py > ${O}_base$i.log 2>&1; echo "base $i exit $?"; grep "cold L2" ${O}_base$i.log
done
for m in 2 3; do VS_C1B_BLOCKS_PER_SM=$m timeout 200 python tools/head_bench.py 2>&1 | grep conv1x1_bwd; done
