#!/bin/bash
# attention dropout: four decisions per hash (drop_keep4) vs the pair form (VS_LIB_PATH = libvitseg_s28.so)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s30
BASE=$PWD/visiontransformer_b200/lib/libvitseg_s28.so
timeout 400 python tools/kernel_probe.py attn > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -1 ${O}_probe.log; grep -v PASS ${O}_probe.log | head -10
timeout 600 python -m pytest tests/test_dropout_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > ${O}_pytest_drop.log 2>&1; echo "pytest dropout exit $?"; tail -3 ${O}_pytest_drop.log
for i in 1 2; do
timeout 200 python tools/attn_bench.py > ${O}_attn_new$i.log 2>&1; echo "new"; grep -E "p=0.1" ${O}_attn_new$i.log
VS_LIB_PATH=$BASE timeout 200 python tools/attn_bench.py > ${O}_attn_base$i.log 2>&1; echo "base"; grep -E "p=0.1" ${O}_attn_base$i.log
done
AB=32 AN=1025 timeout 200 python tools/attn_bench.py 2>&1 | grep -E "fwd p=0.1|bwd p=0.1"
VS_LIB_PATH=$BASE AB=32 AN=1025 timeout 200 python tools/attn_bench.py 2>&1 | grep -E "fwd p=0.1|bwd p=0.1"
timeout 600 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
