#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "component or unique or golden or reference_suite or extremes or device_tensors" 2>&1 | tail -5 > gpurun_out/${tag}_pytest_subset.log
timeout 600 python -m pytest tests/test_configs_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${tag}_pytest_configs.log
timeout 300 python benchmarks/bench_ops.py --ops component_map > gpurun_out/${tag}_k7.jsonl 2> gpurun_out/${tag}_k7.err
timeout 300 python benchmarks/exp_unique_dtypes.py > gpurun_out/${tag}_unique_dtypes.jsonl 2> gpurun_out/${tag}_unique_dtypes.err
FASTREMAP_B200_SO=$PWD/build_variants/libfastremap_b200_FRB_STREAM_STAGES2.so timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except,unique_sparse,component_map > gpurun_out/${tag}_stages2.jsonl 2> gpurun_out/${tag}_stages2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_component_map.csv python benchmarks/prof_ops.py 1024 component_map > /dev/null 2>&1
FRB_PROBE_WINDOW_MIB=48 timeout 300 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_window_probe" -c 2 --csv --log-file gpurun_out/${tag}_ncu_window_probe.csv python benchmarks/exp_window_probe.py inplace > /dev/null 2>&1
