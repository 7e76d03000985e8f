#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
for mode in 0 5; do
  FRB_K3_MODE=$mode timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except > gpurun_out/${tag}_k3_mode$mode.jsonl 2> gpurun_out/${tag}_k3_mode$mode.err
done
FRB_K3_MODE=5 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "remap or mask or lookup or golden or reference_suite or extremes" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
