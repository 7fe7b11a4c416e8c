#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s17
timeout 180 python tools/kernel_probe.py attn > ${O}_probe.log 2>&1; echo "probe exit $?"; grep -c PASS ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
timeout 120 python tools/attn_bench.py > ${O}_attn.log 2>&1; echo "attn bench exit $?"; cat ${O}_attn.log
