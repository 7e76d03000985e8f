#!/bin/bash
# usage: benchmarks/run_multi.sh <tag> <N>  -- parity, weak + strong scaling bench lines and the PCIe floor at N GPUs
tag=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR benchmarks/check_sharded.py > gpurun_out/${tag}_check_n$N.log 2>&1; echo "check rc=$?" >> gpurun_out/${tag}_check_n$N.log
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_weak_n$N.json 2> gpurun_out/${tag}_bench_weak_n$N.err
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/${tag}_stages_weak_n$N.json 2> gpurun_out/${tag}_stages_weak_n$N.err
$TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong > gpurun_out/${tag}_bench_strong_n$N.json 2> gpurun_out/${tag}_bench_strong_n$N.err
$TR benchmarks/pcie_probe_ranks.py > gpurun_out/${tag}_pcie_n$N.jsonl 2> gpurun_out/${tag}_pcie_n$N.err
