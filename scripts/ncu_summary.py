#!/usr/bin/env python
"""dev tool: summarise an .ncu-rep (raw page + source page) for kernels matching a regex.  usage: ncu_summary.py rep [regex]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "k_elastic"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    key = name[:60]
    if key in seen:
        continue
    seen.add(key)
    print('====', name[:110])
    for w in WANT:
        if w in idx:
            print(f'  {w:70s} {r[idx[w]]} {rows[1][idx[w]]}')
    st = [(h, float(r[i])) for h, i in idx.items() if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('_not_issued') and r[i] not in ('', 'n/a')]
    tot = sum(v for _, v in st) or 1
    print('  stalls: ' + ', '.join(f'{h[33:]} {100 * v / tot:.1f}%' for h, v in sorted(st, key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['# Samples']].isdigit()]  # several captured launches repeat the header
    tot = sum(int(r[idx['# Samples']]) for r in data) or 1
    op = collections.Counter(); exe = collections.Counter()
    for r in data:
        s = r[idx['Source']].split()
        o = (s[0] if not s[0].startswith('@') else s[1]).split('.')[0]
        op[o] += int(r[idx['# Samples']]); exe[o] += int(r[idx['Instructions Executed']])
    print('---- source page of', rows[0][1][:80])
    for o, c in op.most_common(10):
        print(f'  {o:10s} samples {100 * c / tot:5.1f}%  executed {exe[o] / 1e6:8.1f}M')
    st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:14]:
        top = sorted(((int(r[idx[h]]), h) for h in st), reverse=True)[:2]
        print('  ', r[idx['# Samples']].rjust(6), r[idx['Source']][:64].ljust(64), top)
