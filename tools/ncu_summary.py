"""summarise an .ncu-rep: per kernel key metrics + top stall reasons + hottest SASS lines"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
idx = [h.index(w) for w in want if w in h]
print(" | ".join(h[i].split('.')[0][-28:] for i in idx))
print(" | ".join(rows[1][i] for i in idx))
for r in rows[2:]:
    print(" | ".join(r[i][:44] for i in idx))
stall_cols = [i for i, n in enumerate(h) if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('_per_issue_active.ratio')] or \
             [i for i, n in enumerate(h) if 'smsp__average_warp_latency_issue_stalled' in n]
for r in rows[2:]:
    st = sorted(((float(r[i] or 0), h[i].replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', '').replace('.ratio', '')) for i in stall_cols), reverse=True)[:5]
    print(r[h.index('Kernel Name')][:30], "stalls:", ", ".join(f"{n}={v:.1f}" for v, n in st))
