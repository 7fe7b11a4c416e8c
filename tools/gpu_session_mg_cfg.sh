#!/bin/bash
# the other BASELINE configs on N GPUs (configs[2], [3], [4])
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out/r02_cfg_n${N}
for c in paed_bin paed_multi vitl384 infer512; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --config $c --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_${c}.json 2> ${O}_${c}.err
  echo "$c exit $?"; python - <<PY
import json
try:
    d=json.loads(open("${O}_${c}.json").read().strip().splitlines()[-1])
    print("  ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(d["roofline"]["frac"],3), d["config"].get("grad_allreduce","")[:12], "batch/gpu", d["config"]["batch_per_gpu"], "ok", (d.get("dp_check") or {}).get("ok"))
except Exception as e:
    print("  no json:", e)
PY
  tail -2 ${O}_${c}.err | cut -c1-200
done
