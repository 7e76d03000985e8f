#!/bin/bash
# last check of the committed code: the whole GPU suite, smoke(), the N=1 bench line
tag=${1:-r02r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench_line.err
