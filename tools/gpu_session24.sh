#!/bin/bash
# head / embedding glue kernels: im2col (warp per pixel-tap), patchify (block per patch row), conv1x1 grids, embed_bwd ILP
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s24
BASE=$PWD/visiontransformer_b200/lib/libvitseg_base.so
timeout 400 python tools/kernel_probe.py head loss gemm > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -1 ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
timeout 300 python tools/step_breakdown.py > ${O}_bd_new.log 2>&1; echo "bd new exit $?"; grep -A40 "all libvitseg" ${O}_bd_new.log | grep -v "gemm$"
VS_LIB_PATH=$BASE timeout 300 python tools/step_breakdown.py > ${O}_bd_base.log 2>&1; echo "bd base exit $?"; grep -A40 "all libvitseg" ${O}_bd_base.log | grep -v "gemm$"
for m in 2 4; do VS_C1B_BLOCKS_PER_SM=$m timeout 300 python tools/step_breakdown.py > ${O}_bd_c1b$m.log 2>&1; echo "c1b $m"; grep "conv1x1_bwd" ${O}_bd_c1b$m.log; done
timeout 600 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
