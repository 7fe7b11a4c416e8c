#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s27
BASE=$PWD/visiontransformer_b200/lib/libvitseg_base.so
timeout 400 python tools/kernel_probe.py head loss > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -1 ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
timeout 200 python tools/head_bench.py > ${O}_new.log 2>&1; echo "new exit $?"; grep -v agreement ${O}_new.log
VS_LIB_PATH=$BASE timeout 200 python tools/head_bench.py > ${O}_base.log 2>&1; echo "base exit $?"; grep "upsample_ce\|upsample_argmax   " ${O}_base.log
