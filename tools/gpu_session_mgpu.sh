#!/bin/bash
# multi-GPU session: usage tools/gpu_session_mgpu.sh N  (run under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02_mg${N}
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_${name}.json 2> ${O}_${name}.err
  echo "$name exit $?"; python - <<PY
import json
try:
    d=json.loads(open("${O}_${name}.json").read().strip().splitlines()[-1])
    print("  ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "gemm frac", round(d["roofline"]["frac"],3), "dp_check", d.get("dp_check",{}).get("ok"), d.get("dp_check"))
except Exception as e:
    print("  no json:", e)
PY
}
run static VS_GEMM_SCHED=static
run clc VS_GEMM_SCHED=clc
run static_nopdl VS_GEMM_SCHED=static VS_PDL=0
run clc_maxcta8 VS_GEMM_SCHED=clc NCCL_MAX_CTAS=8
run clc_maxcta16 VS_GEMM_SCHED=clc NCCL_MAX_CTAS=16
tail -5 ${O}_static.err
echo "== timelines"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_timeline.py --out gpurun_out/r02_timeline_n${N}.csv 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/dp_timeline.py --out gpurun_out/r02_timeline_n1.csv 2>&1 | tail -3
python tools/summarize_timeline.py gpurun_out/r02_timeline_n1.csv gpurun_out/r02_timeline_n${N}.csv | tee gpurun_out/r02_timeline_n${N}_summary.md
VS_GEMM_SCHED=clc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dp_timeline.py --out gpurun_out/r02_timeline_n${N}_clc.csv 2>&1 | tail -3
python tools/summarize_timeline.py gpurun_out/r02_timeline_n1.csv gpurun_out/r02_timeline_n${N}_clc.csv | tee gpurun_out/r02_timeline_n${N}_clc_summary.md
