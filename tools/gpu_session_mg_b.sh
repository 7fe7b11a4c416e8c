#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02_mgb${N}
run() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_${name}.json 2> ${O}_${name}.err
  echo "$name exit $?"; python - <<PY
import json
try:
    d=json.loads(open("${O}_${name}.json").read().strip().splitlines()[-1])
    print("  ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "gemm frac", round(d["roofline"]["frac"],3), "allreduce:", d["config"].get("grad_allreduce","")[:40], d["config"].get("grad_allreduce_fallback"), "dp_check", d.get("dp_check"))
except Exception as e:
    print("  no json:", e)
PY
  tail -3 ${O}_${name}.err | cut -c1-300
}
run multimem VS_DP_REDUCE=multimem
run nccl VS_DP_REDUCE=nccl
run auto
echo "== timeline (multimem)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/dp_timeline.py --out gpurun_out/r02_timeline_n${N}_multimem.csv 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/dp_timeline.py --out gpurun_out/r02_timeline_n1_b.csv 2>&1 | tail -2
python tools/summarize_timeline.py gpurun_out/r02_timeline_n1_b.csv gpurun_out/r02_timeline_n${N}_multimem.csv | tee gpurun_out/r02_timeline_n${N}_multimem_summary.md
