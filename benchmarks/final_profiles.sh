#!/bin/bash
# round-2 evidence run on one B200: tests, both bench arms, the launch list of the bench command, ncu --set full summaries.
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest_gpu.log
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
(time python bench.py --steps 20 --warmup 5) > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench_line.err
python benchmarks/bench_ops.py > gpurun_out/${tag}_ops_1024.jsonl 2> gpurun_out/${tag}_ops_1024.err
python benchmarks/bench_secondary.py > gpurun_out/${tag}_secondary_kernels.jsonl 2> gpurun_out/${tag}_secondary_kernels.err
python benchmarks/bench_layout.py > gpurun_out/${tag}_bench_layout.jsonl 2> gpurun_out/${tag}_bench_layout.err
python benchmarks/exp_unique_dtypes.py > gpurun_out/${tag}_unique_dtypes.jsonl 2> gpurun_out/${tag}_unique_dtypes.err
python benchmarks/exp_renumber_dtypes.py > gpurun_out/${tag}_renumber_dtypes.jsonl 2> gpurun_out/${tag}_renumber_dtypes.err
# launch list of the bench command (device-resident loop only: the e2e / CPU legs launch nothing of interest)
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
# K1 / K3 of the headline: one warm launch each, full set
ncu --set full --clock-control none -k regex:"k_relabel_auto|k_renumber_build" -s 6 -c 2 -f -o gpurun_out/${tag}_k1k3 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_ncu_k1k3.log 2>&1
ncu -i gpurun_out/${tag}_k1k3.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_k1_k3.csv 2>/dev/null
rm -f gpurun_out/${tag}_k1k3.ncu-rep
for op in unique_sparse mask_except remap component_map; do
  case $op in
    unique_sparse) pat="^k_hist$|k_hist_apply_log"; skip=2; cnt=2;;
    mask_except) pat="k_relabel_batched"; skip=1; cnt=1;;
    remap) pat="k_relabel_runs"; skip=1; cnt=1;;
    component_map) pat="k_component_build|k_component_gather_list"; skip=2; cnt=2;;
  esac
  python benchmarks/prof_ops.py 1024 $op > gpurun_out/${tag}_${op}_plain.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"$pat" -s $skip -c $cnt -f -o gpurun_out/${tag}_${op} \
      python benchmarks/prof_ops.py 1024 $op > gpurun_out/${tag}_${op}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_${op}.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_${op}.csv 2>/dev/null
  rm -f gpurun_out/${tag}_${op}.ncu-rep
done
du -sh gpurun_out
