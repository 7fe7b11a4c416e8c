#!/bin/bash
# tail-split GEMM schedule, live fused column sums, attention backward dead-warp skip: parity, then A/B timings
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s13
timeout 600 python tools/kernel_probe.py gemm attn > ${O}_probe.log 2>&1; echo "probe exit $?"; grep -c OK ${O}_probe.log; grep -v OK ${O}_probe.log | head -30
VS_GEMM_TAIL_MN=1 timeout 600 python tools/kernel_probe.py gemm > ${O}_probe_mn.log 2>&1; echo "probe mn exit $?"; grep -c OK ${O}_probe_mn.log; grep -v OK ${O}_probe_mn.log | head -30
VS_GEMM_TAIL=0 timeout 300 python tools/gemm_bench.py model > ${O}_gemm_tail0.log 2>&1; cat ${O}_gemm_tail0.log
timeout 300 python tools/gemm_bench.py model > ${O}_gemm_tail4.log 2>&1; cat ${O}_gemm_tail4.log
VS_GEMM_TAIL_MN=1 timeout 300 python tools/gemm_bench.py model > ${O}_gemm_tail4mn.log 2>&1; cat ${O}_gemm_tail4mn.log
timeout 300 python tools/attn_bench.py > ${O}_attn.log 2>&1; cat ${O}_attn.log
VS_GEMM_TAIL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_tail0.json 2> ${O}_bench_tail0.err; echo "bench tail0 exit $?"; cut -c1-400 ${O}_bench_tail0.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_tail4.json 2> ${O}_bench_tail4.err; echo "bench tail4 exit $?"; cut -c1-400 ${O}_bench_tail4.json
VS_FUSE_COLSUM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_tail4_fuse.json 2> ${O}_bench_tail4_fuse.err; echo "bench fuse exit $?"; cut -c1-400 ${O}_bench_tail4_fuse.json
VS_FUSE_COLSUM=1 VS_GEMM_RASTER=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_tail4_fuse_nor.json 2> ${O}_bench_tail4_fuse_nor.err; echo "bench fuse noraster exit $?"; cut -c1-400 ${O}_bench_tail4_fuse_nor.json
