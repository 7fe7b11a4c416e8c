#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"; cut -c1-300 ${O}_bench.json; tail -2 ${O}_bench.err
timeout 600 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 ${O}_pytest.log
