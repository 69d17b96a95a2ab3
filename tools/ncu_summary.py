#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md/profiles cite."""
import csv, subprocess, sys, json
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct','lts__t_bytes.sum','sm__cycles_elapsed.max',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warp_latency_issue_stalled_barrier.ratio',
 'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio','smsp__average_warp_latency_issue_stalled_mio_throttle.ratio',
 'smsp__average_warp_latency_issue_stalled_lg_throttle.ratio','smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio',
 'smsp__average_warp_latency_issue_stalled_wait.ratio','smsp__average_warp_latency_issue_stalled_not_selected.ratio',
 'smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
def main(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index('Kernel Name')
    res = []
    for r in data:
        d = {'kernel': r[name_i][:60]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k); d[k] = r[i] + ' ' + units[i]
        res.append(d)
    for k in ['kernel']+KEYS:
        vals = [d.get(k,'') for d in res]
        if any(vals): print(f"{k:90s} " + ' | '.join(vals))
if __name__ == '__main__':
    main(sys.argv[1])
