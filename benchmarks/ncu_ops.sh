#!/bin/bash
# ncu --set full captures of the non-headline kernels (one launch each, the second = warm one), with CSV exports.
# usage: benchmarks/ncu_ops.sh <tag> [ops...]   (ops: unique_sparse mask_except component_map remap unique_dense)
tag=$1; shift
ops=${@:-"unique_sparse mask_except component_map"}
mkdir -p gpurun_out
for op in $ops; do
  case $op in
    unique_sparse|unique_dense) pat="^k_hist$"; skip=1;;
    mask_except) pat="k_relabel"; skip=1;;
    remap) pat="k_relabel"; skip=1;;
    component_map) pat="k_component_build"; skip=1;;
  esac
  python benchmarks/prof_ops.py 1024 $op > gpurun_out/${tag}_${op}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o gpurun_out/${tag}_${op} \
      python benchmarks/prof_ops.py 1024 $op > gpurun_out/${tag}_${op}_ncu.log 2>&1
  if [ -f gpurun_out/${tag}_${op}.ncu-rep ]; then
    ncu -i gpurun_out/${tag}_${op}.ncu-rep --page raw --csv > gpurun_out/${tag}_${op}_raw.csv 2>/dev/null
    ncu -i gpurun_out/${tag}_${op}.ncu-rep --page source --csv > gpurun_out/${tag}_${op}_source.csv 2>/dev/null
  fi
done
du -sh gpurun_out
# stay under the 64 MiB return limit: drop the binary reports if needed (the CSV exports carry the numbers)
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/*.ncu-rep; fi
