#!/bin/bash
tag=${1:-r02o}
mkdir -p gpurun_out
V1=$PWD/build_variants/libfastremap_b200_FRB_K3_TABLE_HINT.so
V2=$PWD/build_variants/libfastremap_b200_FRB_K3_TABLE_HINT_FRB_K3_STREAM_EF.so
FASTREMAP_B200_SO=$V1 timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except > gpurun_out/${tag}_table_hint.jsonl 2> gpurun_out/${tag}_table_hint.err
FASTREMAP_B200_SO=$V2 timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except > gpurun_out/${tag}_table_hint_ef.jsonl 2> gpurun_out/${tag}_table_hint_ef.err
