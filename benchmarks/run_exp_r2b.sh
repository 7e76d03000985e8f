#!/bin/bash
# round 2, second session: A/B of the K3 run-head variants, K7 queued offers, packed merged tobytes, the L2 window probe
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tobytes or component or remap or transpos or layout or lookup or golden" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
timeout 600 python -m pytest tests/test_configs_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${tag}_pytest_configs.log
for mode in 0 1 3; do
  FRB_K3_MODE=$mode timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except > gpurun_out/${tag}_k3_mode$mode.jsonl 2> gpurun_out/${tag}_k3_mode$mode.err
done
timeout 300 python benchmarks/bench_ops.py --ops component_map > gpurun_out/${tag}_k7.jsonl 2> gpurun_out/${tag}_k7.err
timeout 300 python benchmarks/bench_layout.py > gpurun_out/${tag}_layout.jsonl 2> gpurun_out/${tag}_layout.err
for mib in 16 32 48 64; do
  FRB_PROBE_WINDOW_MIB=$mib timeout 300 python benchmarks/exp_window_probe.py inplace >> gpurun_out/${tag}_window.jsonl 2>> gpurun_out/${tag}_window.err
done
FRB_PROBE_WINDOW_MIB=32 timeout 300 python benchmarks/exp_window_probe.py >> gpurun_out/${tag}_window.jsonl 2>> gpurun_out/${tag}_window.err
