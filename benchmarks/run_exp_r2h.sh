#!/bin/bash
tag=${1:-r02h}
mkdir -p gpurun_out
FRB_K3_MODE=5 timeout 300 ncu --set full --clock-control none -k regex:"k_relabel_runs" -s 1 -c 1 -f -o gpurun_out/${tag}_runs python benchmarks/prof_ops.py 1024 mask_except > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_runs.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_mask_except_runs.csv 2>/dev/null
rm -f gpurun_out/${tag}_runs.ncu-rep
