#!/bin/bash
tag=${1:-r02t}
mkdir -p gpurun_out
for band in 0 8 4 16; do
  FRB_PLANE_BAND=$band timeout 300 python benchmarks/bench_ops.py --ops unique_sparse,unique_dense,component_map > gpurun_out/${tag}_band$band.jsonl 2> gpurun_out/${tag}_band$band.err
done
timeout 600 python -m pytest tests/test_configs_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "2a or 2b or 5b or full_size or row_geometries or listed_slots" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
