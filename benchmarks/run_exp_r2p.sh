#!/bin/bash
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tobytes or transpos or layout or listed_slots or fortran" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
timeout 300 python benchmarks/bench_layout.py > gpurun_out/${tag}_bench_layout.jsonl 2> gpurun_out/${tag}_bench_layout.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_stages_n1.json 2> gpurun_out/${tag}_stages_n1.err
