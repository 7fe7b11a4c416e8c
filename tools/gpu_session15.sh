#!/bin/bash
# ncu --set full (+source) of the short-sequence attention backward
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s15
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_bwd_short -s 30 -c 1 -o ${O}_attn_bwd_short -f python tools/attn_bench.py > ${O}_ncu.log 2>&1; echo "ncu exit $?"; tail -3 ${O}_ncu.log
ncu -i ${O}_attn_bwd_short.ncu-rep --page raw --csv > ${O}_attn_bwd_short_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep s15
