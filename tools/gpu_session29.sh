#!/bin/bash
# the other BASELINE configs on one GPU with the end-of-round build
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s29
for c in paed_bin paed_multi vitl384 infer512; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_${c}.json 2> ${O}_${c}.err
  echo "$c exit $?"; python - <<PY
import json
try:
    d=json.loads(open("${O}_${c}.json").read().strip().splitlines()[-1])
    print("  ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(d["roofline"]["frac"],3), "TF", round(d["roofline"]["achieved"]), d.get("logits_images_per_sec"))
except Exception as e:
    print("  no json:", e)
PY
  tail -2 ${O}_${c}.err | cut -c1-200
done
