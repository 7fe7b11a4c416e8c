#!/bin/bash
# short-sequence attention backward (whole (batch, head) items per CTA): parity, micro-benchmark, step A/B
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s14
timeout 180 python tools/kernel_probe.py attn > ${O}_probe.log 2>&1; echo "probe exit $?"; grep -c PASS ${O}_probe.log; grep -v PASS ${O}_probe.log | head -30
timeout 180 python -m pytest tests/test_dropout_gpu.py -x -q -m gpu > ${O}_dropout.log 2>&1; echo "dropout test exit $?"; tail -5 ${O}_dropout.log
timeout 120 python tools/attn_bench.py > ${O}_attn.log 2>&1; echo "attn bench exit $?"; cat ${O}_attn.log
VS_ATTN_BWD=blocks timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_blocks.json 2> ${O}_bench_blocks.err; echo "bench blocks exit $?"; cut -c1-300 ${O}_bench_blocks.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_short.json 2> ${O}_bench_short.err; echo "bench short exit $?"; cut -c1-300 ${O}_bench_short.json; tail -3 ${O}_bench_short.err
