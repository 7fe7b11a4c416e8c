#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02_mgd${N}
run() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_${name}.json 2> ${O}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("${O}_${name}.json").read().strip().splitlines()[-1])
    print("$name: ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "gemm frac", round(d["roofline"]["frac"],3), d["config"].get("grad_allreduce","")[:12], "ok", d.get("dp_check",{}).get("ok"))
except Exception as e:
    print("$name: no json:", e)
PY
}
run c148 VS_MM_CTAS=148
run c64 VS_MM_CTAS=64
run c32 VS_MM_CTAS=32
run c16 VS_MM_CTAS=16
run c8 VS_MM_CTAS=8
run c32carve VS_MM_CTAS=32 VS_MM_CARVEOUT=1
run c16carve VS_MM_CTAS=16 VS_MM_CARVEOUT=1
run c148b VS_MM_CTAS=148
