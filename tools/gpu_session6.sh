#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s6
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > ${O}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 ${O}_pytest.log | cut -c1-300
for pv in 0 4 2; do echo "VS_ATTN_POLY=$pv"; AB=32 AN=1025 AH=12 VS_ATTN_POLY=$pv timeout 200 python tools/attn_bench.py 2>&1 | grep -E "fwd"; done
for pv in 0 4; do echo "N=577 VS_ATTN_POLY=$pv"; AB=32 AN=577 AH=16 VS_ATTN_POLY=$pv timeout 200 python tools/attn_bench.py 2>&1 | grep -E "fwd"; done
for pv in 0 4 2; do VS_ATTN_POLY=$pv timeout 900 python bench.py --config infer512 --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline > ${O}_infer512_poly${pv}.json 2> ${O}_infer512_poly${pv}.err; echo "infer512 poly$pv: $(cut -c1-170 ${O}_infer512_poly${pv}.json)"; done
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"; cut -c1-200 ${O}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "ref exit $?"; cut -c1-300 ${O}_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
