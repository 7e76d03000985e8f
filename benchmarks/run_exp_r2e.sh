#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "component or golden or reference_suite or extremes or device_tensors" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
timeout 600 python -m pytest tests/test_configs_gpu.py -m gpu -q -x -k "5b or component" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_configs.log
timeout 300 python benchmarks/bench_ops.py --ops component_map > gpurun_out/${tag}_k7.jsonl 2> gpurun_out/${tag}_k7.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_component_map.csv python benchmarks/prof_ops.py 1024 component_map > /dev/null 2>&1
