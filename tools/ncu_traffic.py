"""profiles/rNN_traffic.json from an .ncu-rep (ncu --set full): per kernel (first launch of each name in the report)
DRAM bytes, duration, registers, occupancy, issue utilisation, executed instructions.
    python tools/ncu_traffic.py REP OUT.json "source note" [raw.csv]"""
import csv, io, json, subprocess, sys
rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
if len(sys.argv) > 4:
    open(sys.argv[4], "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}


def val(r, name, scale_to=None):
    if name not in col or r[col[name]] == "":
        return None
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    if scale_to == "byte":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if scale_to == "ms":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
    return v


kernels = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    if name in kernels:
        continue
    kernels[name] = {
        "dram_bytes_per_launch": (val(r, "dram__bytes_read.sum", "byte") or 0) + (val(r, "dram__bytes_write.sum", "byte") or 0),
        "gpu_time_ms": val(r, "gpu__time_duration.sum", "ms"),
        "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers": val(r, "launch__registers_per_thread"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": val(r, "smsp__inst_executed.sum"),
    }
import os
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
if not commit and os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".commit_id")):   # the GPU box has no .git
    commit = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".commit_id")).read().strip()
json.dump({"source": note, "commit": commit, "kernels": kernels}, open(out, "w"), indent=1)
print(out, len(kernels), "kernels")
