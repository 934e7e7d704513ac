#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + per-source-line instruction shares.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_warps_for_per_warp_column]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_adu.sum',
        'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_xu.sum']
for vals in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if h in KEEP or 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
            print(f"{h:80s} {units[i]:>16s} {vals[i]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
per, tot, cur_file, hdr2 = {}, 0.0, None, None
for r in csv.reader(src.splitlines()):
    if r and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr2 = r
        ie, smp = hdr2.index('Instructions Executed'), hdr2.index('# Samples')
        continue
    if hdr2 is None or len(r) <= ie or r[0] == '' or r[2] != '-':
        continue
    try:
        n, s = float(r[ie]), float(r[smp])
    except ValueError:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:95])
    a = per.get(key, (0, 0))
    per[key] = (a[0] + n, a[1] + s)
    tot += n
nw = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("=" * 100)
print(f"total warp instructions {tot:.0f}" + (f" = {tot / nw:.1f} per unit" if nw else ""))
stot = sum(v[1] for v in per.values())
for k, (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{n / tot * 100:6.2f}% inst {s / max(stot, 1) * 100:6.2f}% smp " + (f"{n / nw:8.1f} " if nw else "") + f"{k[0]}:{k[1]}  {k[2]}")
