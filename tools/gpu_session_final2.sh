#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_final
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > ${O}_n2.json 2> ${O}_n2.err; echo "n2 exit $?"; cut -c1-900 ${O}_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_n1.json 2> ${O}_n1.err; echo "n1 exit $?"; cut -c1-2500 ${O}_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > ${O}_ref_n2.json 2> ${O}_ref_n2.err; echo "ref n2 exit $?"; cut -c1-200 ${O}_ref_n2.json
