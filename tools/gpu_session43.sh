#!/bin/bash
# last sanity of the other configs with the final library (few steps)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s43
for c in paed_bin vitl384 infer512; do
  timeout 300 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_${c}.json 2> ${O}_${c}.err
  echo "$c exit $?"; python - <<PY
import json
try:
    d=json.loads(open("${O}_${c}.json").read().strip().splitlines()[-1])
    print("  ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "final", d.get("final_result"))
except Exception as e:
    print("  no json:", e)
PY
  tail -1 ${O}_${c}.err | cut -c1-200
done
