#!/bin/bash
# usage: benchmarks/run_multi_short.sh <tag> <N>  -- weak + strong scaling bench lines (with parity) and the stage split at N GPUs
tag=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_weak_n$N.json 2> gpurun_out/${tag}_bench_weak_n$N.err
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/${tag}_stages_weak_n$N.json 2> gpurun_out/${tag}_stages_weak_n$N.err
$TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-e2e > gpurun_out/${tag}_stages_strong_n$N.json 2> gpurun_out/${tag}_stages_strong_n$N.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_stages_n1.json 2> gpurun_out/${tag}_stages_n1.err
