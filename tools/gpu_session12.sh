#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_s12
timeout 120 python tools/ncu_one_gemm.py 12608 3072 768 gelu out2 cfg=1 2>&1 | tail -1 | cut -c1-120
timeout 120 python tools/ncu_one_gemm.py 12608 3072 768 aux1 bmn cfg=1 2>&1 | tail -1 | cut -c1-120
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -s 4 -c 1 -o ${O}_gemm_fc1fwd -f python tools/ncu_one_gemm.py 12608 3072 768 gelu out2 cfg=1 > ${O}_ncu_a.log 2>&1; echo "ncu fc1 exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -s 4 -c 1 -o ${O}_gemm_fc2dgrad -f python tools/ncu_one_gemm.py 12608 3072 768 aux1 bmn cfg=1 > ${O}_ncu_b.log 2>&1; echo "ncu fc2dgrad exit $?"
for n in fc1fwd fc2dgrad; do ncu -i ${O}_gemm_${n}.ncu-rep --page raw --csv > ${O}_gemm_${n}_raw.csv 2>/dev/null; done
python - <<'PY'
import csv
for n in ("fc1fwd","fc2dgrad"):
    rows=list(csv.reader(open(f"gpurun_out/r02_s12_gemm_{n}_raw.csv")))
    hdr=rows[0]
    want=['gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__registers_per_thread']
    r=rows[2]
    print(n, {w:r[hdr.index(w)] for w in want if w in hdr})
PY
