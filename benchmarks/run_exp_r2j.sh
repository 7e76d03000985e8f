#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 600 python benchmarks/exp_k3_variants.py > gpurun_out/${tag}_k3_variants.jsonl 2> gpurun_out/${tag}_k3_variants.err
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "remap or mask or lookup" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
