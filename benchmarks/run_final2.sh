#!/bin/bash
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest_gpu.log
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench_line.err
timeout 400 python benchmarks/bench_ops.py > gpurun_out/${tag}_ops_1024.jsonl 2> gpurun_out/${tag}_ops_1024.err
timeout 300 python benchmarks/exp_unique_dtypes.py > gpurun_out/${tag}_unique_dtypes.jsonl 2> gpurun_out/${tag}_unique_dtypes.err
for op in unique_sparse component_map; do
  case $op in
    unique_sparse) pat="^k_hist$|k_hist_apply_log"; skip=2; cnt=2;;
    component_map) pat="k_component_build_z|k_component_gather_list"; skip=2; cnt=2;;
  esac
  timeout 300 ncu --set full --clock-control none -k regex:"$pat" -s $skip -c $cnt -f -o gpurun_out/${tag}_${op} \
      python benchmarks/prof_ops.py 1024 $op > gpurun_out/${tag}_${op}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_${op}.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_${op}.csv 2>/dev/null
  rm -f gpurun_out/${tag}_${op}.ncu-rep
done
