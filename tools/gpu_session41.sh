#!/bin/bash
# QKV bias gradient formed inside the short attention backward vs the separate column-sum pass (VS_ATTN_BIAS=separate)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s41
timeout 300 python tools/kernel_probe.py attn > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -1 ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
timeout 120 python tools/attn_bench.py > ${O}_attn.log 2>&1; grep -E "short" ${O}_attn.log
timeout 600 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_fused.json 2> ${O}_bench_fused.err; echo "bench fused exit $?"; cut -c1-230 ${O}_bench_fused.json
VS_ATTN_BIAS=separate timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_sep.json 2> ${O}_bench_sep.err; echo "bench separate exit $?"; cut -c1-230 ${O}_bench_sep.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_fused2.json 2> ${O}_bench_fused2.err; echo "bench fused exit $?"; cut -c1-230 ${O}_bench_fused2.json
