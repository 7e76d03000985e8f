python benchmarks/prof_ops.py 1024 unique_sparse > gpurun_out/r02i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hist_apply_log -s 1 -c 1 -f -o gpurun_out/r02i_apply python benchmarks/prof_ops.py 1024 unique_sparse > gpurun_out/r02i_ncu.log 2>&1
ncu -i gpurun_out/r02i_apply.ncu-rep --page raw --csv > gpurun_out/r02i_apply_raw.csv 2>/dev/null
ncu -i gpurun_out/r02i_apply.ncu-rep --page source --csv > gpurun_out/r02i_apply_source.csv 2>/dev/null
rm -f gpurun_out/r02i_apply.ncu-rep
