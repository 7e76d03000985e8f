#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports: the handful of counters the roofline argument needs, per kernel launch.
  python benchmarks/ncu_summary.py gpurun_out/x_raw.csv [...]"""
import csv
import sys

KEYS = [
  "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
  "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
  "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
  "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
  "smsp__inst_executed.sum", "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_atom.sum",
  "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
  "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
  "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
  "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
  "smsp__thread_inst_executed_per_inst_executed.ratio",
  "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
  "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
  for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
      d = dict(zip(hdr, r))
      u = dict(zip(hdr, units))
      print("== %s :: %s" % (path, d.get("Kernel Name", "?")[:110]))
      for k in KEYS:
        if k in d:
          print("  %-62s %s %s" % (k, d[k], u[k]))
      stalls = sorted(((float(d[k]), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k in hdr
                       if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k]), reverse=True)
      print("  stalls (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))


if __name__ == "__main__":
  main()
