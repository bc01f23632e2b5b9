#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one row per distinct kernel."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64%'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'blk'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'bar'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'short'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'wait'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'mio'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'lg'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'math'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf'), ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smemwf'),
        ('lts__t_sector_hit_rate.pct', 'l2hit%'), ('smsp__inst_executed.sum', 'inst')]
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
seen = set()
for r in data:
    name = re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '').strip()
    key = name + r[idx['launch__grid_size']]
    if name in seen:
        continue
    seen.add(name)
    vals = []
    for c, _ in cols:
        if c not in idx:
            vals.append('-'); continue
        v = r[idx[c]]; u = units[idx[c]]
        try:
            x = float(v.replace(',', ''))
            if 'byte' in u:
                scale = {'Gbyte': 1e3, 'Mbyte': 1.0, 'Kbyte': 1e-3, 'byte': 1e-6}.get(u, 1.0)
                vals.append(f"{x * scale:.0f}MB")
            else:
                vals.append(f"{x:.3g}")
        except ValueError:
            vals.append(v)
    print(f"| `{name}` | " + " | ".join(vals) + " |")
