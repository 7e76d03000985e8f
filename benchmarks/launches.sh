#!/bin/bash
# per-launch device times of one op (cold-cache, serialised: compare shares).  usage: benchmarks/launches.sh <tag> <op>
tag=$1; op=$2
python benchmarks/prof_ops.py 1024 $op > gpurun_out/${tag}_${op}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_${op}_launches.csv python benchmarks/prof_ops.py 1024 $op > /dev/null 2>&1
