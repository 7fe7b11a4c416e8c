#!/bin/bash
# N-GPU sanity of the end-of-round build: bench under torchrun (dp_check + in-switch reduction)
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02_s31_n${N}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench.json 2> ${O}_bench.err; echo "bench n=$N exit $?"
python - <<PY
import json
d=json.loads(open("${O}_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(d["roofline"]["frac"],3), d["config"].get("grad_allreduce",""), "dp_check", d.get("dp_check"))
PY
tail -3 ${O}_bench.err | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench n=1 exit $?"; cut -c1-200 ${O}_bench_n1.json
