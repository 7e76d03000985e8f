#!/bin/bash
# checkpoint: full GPU suite, the N=1 bench line, the device-resident op lines
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest_gpu.log
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench_line.err
timeout 400 python benchmarks/bench_ops.py > gpurun_out/${tag}_ops_1024.jsonl 2> gpurun_out/${tag}_ops_1024.err
timeout 300 python benchmarks/exp_unique_dtypes.py > gpurun_out/${tag}_unique_dtypes.jsonl 2> gpurun_out/${tag}_unique_dtypes.err
