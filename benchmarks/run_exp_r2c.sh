#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
FRB_K3_MODE=4 timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except > gpurun_out/${tag}_k3_mode4.jsonl 2> gpurun_out/${tag}_k3_mode4.err
FASTREMAP_B200_SO=$PWD/build_variants/libfastremap_b200_FRB_STREAM_STAGES4.so timeout 300 python benchmarks/bench_ops.py --ops remap,mask_except,unique_sparse,unique_dense,component_map > gpurun_out/${tag}_stages4.jsonl 2> gpurun_out/${tag}_stages4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_component_map.csv python benchmarks/prof_ops.py 1024 component_map > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_unique_u8_u16.csv python benchmarks/exp_unique_dtypes.py uint8 uint16 > /dev/null 2>&1
FRB_PROBE_WINDOW_MIB=48 timeout 300 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_window_probe|k_stream_probe" -c 6 --csv --log-file gpurun_out/${tag}_ncu_window_probe.csv python benchmarks/exp_window_probe.py inplace > /dev/null 2>&1
