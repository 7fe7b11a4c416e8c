#!/bin/bash
set -u
mkdir -p gpurun_out
VS_LIB_PATH=$PWD/visiontransformer_b200/lib/libvitseg_trace.so timeout 120 python tools/attn_trace.py 0.1 > gpurun_out/r02_s19_trace.log 2>&1; echo "trace exit $?"; head -5 gpurun_out/r02_s19_trace.log
