#!/bin/bash
tag=${1:-r02w2}
mkdir -p gpurun_out
for r in 16 32 128 256; do
  timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except --tile-run $r > gpurun_out/${tag}_run$r.jsonl 2> gpurun_out/${tag}_run$r.err
done
