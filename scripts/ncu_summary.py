"""Summarise ncu outputs: launch-list CSV (shares per kernel) and a --set full report (key metrics per launch)."""
import csv, subprocess, sys
from collections import defaultdict

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rows[1:]:
        n = r[ki].split('(')[0][-40:]; v = float(r[vi].replace(',', '')); tot[n] += v; cnt[n] += 1
    s = sum(tot.values())
    for n, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{n:42s} {cnt[n]:4d} launches {v/1e6:10.3f} ms {100*v/s:5.1f}%")

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed.sum', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']

def report(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]; units = rows[1]
    ki = h.index('Kernel Name')
    for r in rows[2:]:
        print('==', r[ki].split('(')[0][-40:])
        for i, n in enumerate(h):
            if n in KEYS:
                print(f"   {n:82s} {r[i]:>16s} {units[i]}")

if __name__ == '__main__':
    for p in sys.argv[1:]:
        (launches if p.endswith('.csv') else report)(p)
