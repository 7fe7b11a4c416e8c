#!/bin/bash
# short attention forward: persistent CTAs vs one CTA per item
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s33
timeout 300 python tools/kernel_probe.py attn > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -1 ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
for i in 1 2; do
timeout 120 python tools/attn_bench.py > ${O}_persist$i.log 2>&1; echo "persist"; grep -E "^fwd" ${O}_persist$i.log
VS_ATTN_FWD_SHORT=oneshot timeout 120 python tools/attn_bench.py > ${O}_oneshot$i.log 2>&1; echo "oneshot"; grep -E "^fwd" ${O}_oneshot$i.log
done
