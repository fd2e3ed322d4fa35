import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f pages/s  %.3f ms/step | e2e %.0f pages/s %.3f ms/step | launches %d | n_gpus %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["n_gpus"]))
for k, v in (d.get("e2e_variants") or {}).items():
    print("  e2e variant %-18s %8.0f pages/s  %.3f ms/step" % (k, v["value"], v["ms_per_step"]))
if d.get("with_forward"):
    print("  with_forward %.0f pages/s (forward share %.2f)" % (d["with_forward"]["value"], d["with_forward"]["forward_share_of_step"]))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:16]:
    print("  %-34s %7.3f ms/step  %6.0f GB/s  frac %.2f" % (k, v["ms_per_step"], v.get("gbs", 0) or 0, v.get("frac_of_hbm_peak", 0) or 0))
for k, v in d["summary"].items():
    if "frac_of_hbm_peak" in v:
        print("  %-34s %7.3f ms/step  frac %.2f" % (k, v["ms_per_step"], v["frac_of_hbm_peak"] or 0))
print("  kernel sum %.3f ms" % d["summary"]["whole_path"]["kernel_ms_per_step"])
print("  roofline:", d["roofline"]["kernel"], d["roofline"]["frac"], "clocks", d["clocks"])
if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"])
