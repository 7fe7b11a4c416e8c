#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02_mgc${N}
run() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_${name}.json 2> ${O}_${name}.err
  echo "$name exit $?"; python - <<PY
import json
try:
    d=json.loads(open("${O}_${name}.json").read().strip().splitlines()[-1])
    print("  ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "gemm frac", round(d["roofline"]["frac"],3), "allreduce:", d["config"].get("grad_allreduce","")[:30], d["config"].get("grad_allreduce_fallback"), "dp_check ok", d.get("dp_check",{}).get("ok"), "werr", d.get("dp_check",{}).get("weight_err"))
except Exception as e:
    print("  no json:", e)
PY
}
run auto
run nccl VS_DP_REDUCE=nccl
echo "== timeline (auto)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/dp_timeline.py --out gpurun_out/r02_timeline_n${N}_mm2.csv 2>&1 | grep "rank 0"
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/dp_timeline.py --out gpurun_out/r02_timeline_n1_d.csv 2>&1 | grep "rank 0"
python tools/summarize_timeline.py gpurun_out/r02_timeline_n1_d.csv gpurun_out/r02_timeline_n${N}_mm2.csv | tee gpurun_out/r02_timeline_n${N}_mm2_summary.md | head -8
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_n1.json 2> ${O}_n1.err; echo "n1: $(cut -c1-160 ${O}_n1.json)"
