#!/bin/bash
# Key metrics of an .ncu-rep (run in the dev container): tools/ncu_read.sh file.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2:]
keys=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_bytes.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","launch__grid_size","launch__block_size","smsp__inst_executed.sum","sm__inst_executed.avg.per_cycle_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","sm__cycles_elapsed.max","smsp__cycles_active.avg","launch__waves_per_multiprocessor","sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active","smsp__inst_executed_pipe_xu.sum"]
for v in vals:
  for k in keys:
    for i,h in enumerate(hdr):
      if h==k: print(f"{k:75s} {v[i]:>20s} {units[i]}")
  # stall reasons
  st=[(float(v[i].replace(",","") or 0),h) for i,h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio")]
  for x,h in sorted(st,reverse=True)[:8]: print(f"   stall {h:70s} {x:8.2f}")
'
