#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s7
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > ${O}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|FAILED|ViT-B/16 @512|ViT-L/16 @384|p8/1024/16h P" ${O}_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"; cut -c1-200 ${O}_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
