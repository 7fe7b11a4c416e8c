#!/bin/bash
# round-2 GPU session 1: parity at headline size, new epilogue A/B, first bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_s1_smi.txt 2>&1
nproc > gpurun_out/r02_s1_nproc.txt
echo "== pytest gpu" 
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r02_s1_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r02_s1_pytest.log
tail -40 gpurun_out/r02_s1_pytest.log | cut -c1-300
echo "== epilogue A/B"
VS_GEMM_EPI=direct timeout 300 python tools/gemm_epi_bench.py > gpurun_out/r02_s1_epi_direct.log 2>&1
VS_GEMM_EPI=tma timeout 300 python tools/gemm_epi_bench.py > gpurun_out/r02_s1_epi_tma.log 2>&1
paste -d'|' <(cut -c1-70 gpurun_out/r02_s1_epi_direct.log) <(cut -c45-70 gpurun_out/r02_s1_epi_tma.log) | head -70
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_s1_bench.json 2> gpurun_out/r02_s1_bench.err
echo "bench exit $?"; cut -c1-1500 gpurun_out/r02_s1_bench.json
VS_GEMM_EPI=direct timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/r02_s1_bench_direct.json 2> gpurun_out/r02_s1_bench_direct.err
echo "bench(direct) exit $?"; cut -c1-400 gpurun_out/r02_s1_bench_direct.json
echo "== CLC scheduler"
VS_GEMM_SCHED=clc timeout 600 python tools/kernel_probe.py gemm > gpurun_out/r02_s1_probe_clc.log 2>&1
echo "probe(clc) exit $?"; grep -c PASS gpurun_out/r02_s1_probe_clc.log; grep FAIL gpurun_out/r02_s1_probe_clc.log | head -20
VS_GEMM_SCHED=clc timeout 300 python tools/gemm_bench.py > gpurun_out/r02_s1_gemm_clc.log 2>&1
timeout 300 python tools/gemm_bench.py > gpurun_out/r02_s1_gemm_static.log 2>&1
paste -d'\n' gpurun_out/r02_s1_gemm_static.log gpurun_out/r02_s1_gemm_clc.log | cut -c1-250
VS_GEMM_SCHED=clc timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/r02_s1_bench_clc.json 2> gpurun_out/r02_s1_bench_clc.err
echo "bench(clc) exit $?"; cut -c1-400 gpurun_out/r02_s1_bench_clc.json
