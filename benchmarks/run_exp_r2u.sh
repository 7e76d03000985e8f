#!/bin/bash
tag=${1:-r02u}
mkdir -p gpurun_out
for band in 16 32 64; do
  FRB_PLANE_BAND=$band timeout 300 python benchmarks/bench_ops.py --ops unique_sparse,unique_dense,component_map > gpurun_out/${tag}_band$band.jsonl 2> gpurun_out/${tag}_band$band.err
done
