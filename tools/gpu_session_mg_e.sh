#!/bin/bash
set -u
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out/r02_mge${N}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 > ${O}_ce.json 2> ${O}_ce.err
echo "ce exit $?"; python - <<PY
import json
d=json.loads(open("${O}_ce.json").read().strip().splitlines()[-1])
print("  ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "gemm frac", round(d["roofline"]["frac"],3), d["config"].get("grad_allreduce","")[:12], d.get("dp_check"))
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_n1.json 2> ${O}_n1.err; echo "n1: $(cut -c1-160 ${O}_n1.json)"
bash tools/gpu_session_mg_cfg.sh $N
