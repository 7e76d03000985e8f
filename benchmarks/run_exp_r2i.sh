#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
V1=$PWD/build_variants/libfastremap_b200_FRB_EF_MIN_VECS16777216ull.so
V2=$PWD/build_variants/libfastremap_b200_FRB_EF_MIN_VECS16777216ull_FRB_EF_STORES.so
FASTREMAP_B200_SO=$V1 timeout 400 python benchmarks/bench_ops.py > gpurun_out/${tag}_ef_loads.jsonl 2> gpurun_out/${tag}_ef_loads.err
FASTREMAP_B200_SO=$V2 timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except > gpurun_out/${tag}_ef_loads_stores.jsonl 2> gpurun_out/${tag}_ef_loads_stores.err
FASTREMAP_B200_SO=$V2 FRB_K3_MODE=5 timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except > gpurun_out/${tag}_ef_loads_stores_mode5.jsonl 2> gpurun_out/${tag}_ef_loads_stores_mode5.err
FASTREMAP_B200_SO=$V1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_bench_ef_loads.json 2> gpurun_out/${tag}_bench_ef_loads.err
