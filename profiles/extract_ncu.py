"""Condenses an `ncu --set full` report (read with `ncu -i X.ncu-rep --page raw --csv > raw.csv`) into the
handful of numbers DESIGN.md quotes: one block per distinct (kernel, grid)."""
import csv, sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg',
    'sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
]

def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ni, gi = hdr.index('Kernel Name'), hdr.index('launch__grid_size')
    seen = set()
    for r in data:
        key = (r[ni].split('(')[0], r[gi])
        if key in seen:
            continue
        seen.add(key)
        print('===== %s   grid %s' % key)
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('  %-92s %s %s' % (w, r[i], units[i]))

if __name__ == '__main__':
    main(sys.argv[1])
