"""Key metrics + stall reasons from an ncu report (ncu -i X --page raw --csv piped in via file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct']
for d in data[: int(sys.argv[2]) if len(sys.argv) > 2 else 1]:
    print('kernel:', d[hdr.index('Kernel Name')][:80])
    for w in want:
        if w in hdr: print(f'  {w:75s} {d[hdr.index(w)]} {units[hdr.index(w)]}')
    stall = [(h, d[i]) for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    print('  stalls/issue:', ', '.join(f"{h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')} {float(v):.2f}"
                                      for h, v in sorted(stall, key=lambda t: -float(t[1] or 0))[:9]))
