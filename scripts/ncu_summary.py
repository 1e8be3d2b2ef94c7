#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed): key raw metrics + instruction/stall share per SASS region.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [bucket]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {vals[i]:>20s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); ismp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot_ex = sum(int(r[iex]) for r in data); tot_s = sum(int(r[ismp]) for r in data)
print(f"total warp instructions {tot_ex:.4g}, samples {tot_s}")
tot_st = collections.Counter()
for r in data:
    for i in stall_cols:
        tot_st[hdr[i]] += int(r[i] or 0)
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot_s:.1f}%" for k, v in tot_st.most_common(9)))
for b in range(0, len(data), B):
    blk = data[b:b + B]
    ex = sum(int(r[iex]) for r in blk); s = sum(int(r[ismp]) for r in blk)
    if ex < tot_ex * 0.003 and s < tot_s * 0.003:
        continue
    st = collections.Counter()
    for r in blk:
        for i in stall_cols:
            st[hdr[i]] += int(r[i] or 0)
    ops = collections.Counter((r[isrc].split()[1] if r[isrc].strip().startswith('@') else r[isrc].split()[0]) for r in blk)
    flags = [o for o in ('ATOMS.CAST.SPIN', 'MUFU.RCP', 'BAR.SYNC.DEFER_BLOCKING', 'VOTE.ANY', 'STG.E', 'LDG.E.64.CONSTANT', 'LDS.128', 'SHFL.UP') if o in ops]
    top = ", ".join(f"{k[6:]} {100*v/max(s,1):.0f}%" for k, v in st.most_common(3))
    print(f"{b:5d}-{b+B:5d} inst {100*ex/tot_ex:6.2f}%  samples {100*s/tot_s:6.2f}%  [{top}]  {flags}")
