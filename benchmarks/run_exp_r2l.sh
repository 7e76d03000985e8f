#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "unique or golden or reference_suite or ragged or extremes or inverse or point_cloud" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
timeout 300 python benchmarks/exp_unique_dtypes.py uint8 uint16 > gpurun_out/${tag}_unique_dtypes.jsonl 2> gpurun_out/${tag}_unique_dtypes.err
timeout 300 python benchmarks/exp_small_latency.py > gpurun_out/${tag}_small.jsonl 2> gpurun_out/${tag}_small.err
