#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s18
timeout 120 python tools/attn_bench.py > ${O}_attn_hint.log 2>&1; echo "attn bench (hint) exit $?"; cat ${O}_attn_hint.log
VS_LIB_PATH=$PWD/visiontransformer_b200/lib/libvitseg_nohint.so timeout 120 python tools/attn_bench.py > ${O}_attn_nohint.log 2>&1; echo "attn bench (no hint) exit $?"; cat ${O}_attn_nohint.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_hint.json 2> ${O}_bench_hint.err; echo "bench hint exit $?"; cut -c1-300 ${O}_bench_hint.json
VS_LIB_PATH=$PWD/visiontransformer_b200/lib/libvitseg_nohint.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_nohint.json 2> ${O}_bench_nohint.err; echo "bench nohint exit $?"; cut -c1-300 ${O}_bench_nohint.json
