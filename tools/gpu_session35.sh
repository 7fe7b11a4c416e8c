#!/bin/bash
# validation of the end-of-round build: GPU tests, smoke, kernel probe, default bench (with cpu / library baselines), reference arm
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s35
timeout 600 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 ${O}_smoke.log
timeout 600 python tools/kernel_probe.py > ${O}_probe.log 2>&1; echo "probe exit $?"; tail -1 ${O}_probe.log
timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"; cat ${O}_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "ref exit $?"; cat ${O}_bench_ref.json
timeout 300 python tools/attn_bench.py > ${O}_attn.log 2>&1; grep -E "fwd|bwd" ${O}_attn.log
