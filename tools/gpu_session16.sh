#!/bin/bash
# mbarrier try_wait with a suspend-time hint vs without (VS_LIB_PATH = the no-hint build); new MMA issue order of the short attention backward
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s16
timeout 180 python tools/kernel_probe.py attn > ${O}_probe.log 2>&1; echo "probe exit $?"; grep -c PASS ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
timeout 120 python tools/attn_bench.py > ${O}_attn_hint.log 2>&1; echo "attn bench (hint) exit $?"; cat ${O}_attn_hint.log
VS_LIB_PATH=$PWD/visiontransformer_b200/lib/libvitseg_nohint.so timeout 120 python tools/attn_bench.py > ${O}_attn_nohint.log 2>&1; echo "attn bench (no hint, old MMA order) exit $?"; cat ${O}_attn_nohint.log
timeout 300 python tools/gemm_bench.py model > ${O}_gemm_hint.log 2>&1; cat ${O}_gemm_hint.log
VS_LIB_PATH=$PWD/visiontransformer_b200/lib/libvitseg_nohint.so timeout 300 python tools/gemm_bench.py model > ${O}_gemm_nohint.log 2>&1; cat ${O}_gemm_nohint.log
