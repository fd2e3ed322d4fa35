"""per-source-line instruction / stall-sample totals of one launch of an .ncu-rep (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py REP LAUNCH_INDEX [TOP]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, data, h = "", [], None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        print(r[1][:100]); continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        h = r; ie = h.index("Instructions Executed"); ss = h.index("# Samples"); continue
    if h and r[0].isdigit() and r[ie].isdigit():
        data.append((int(r[ie]), int(r[ss]), fname, int(r[0]), r[1]))
tot, tots = sum(d[0] for d in data), sum(d[1] for d in data)
print("total warp instructions", tot, "samples", tots)
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]:>10} {100 * d[0] / tot:5.1f}%  smp {100 * d[1] / max(tots, 1):5.1f}%  {d[2]}:{d[3]}  {d[4].strip()[:110]}")
