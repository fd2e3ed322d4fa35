import json, sys
d = json.load(open(sys.argv[1]))
print("value %.0f pages/s  %.3f ms/step | e2e %.0f pages/s %.3f ms/step | launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]:
    print("  %-34s %7.3f ms/step  %6.0f GB/s  frac %.2f" % (k, v["ms_per_step"], v.get("gbs", 0) or 0, v.get("frac_of_hbm_peak", 0) or 0))
print("  kernel sum %.3f ms, db unit %.3f ms (%.0f GB/s)" % (d["summary"]["whole_path"]["kernel_ms_per_step"], d["summary"]["db_postprocess_unit"]["ms_per_step"], d["summary"]["db_postprocess_unit"]["gbs"]))
print("  roofline:", d["roofline"]["kernel"], d["roofline"]["frac"], "clocks", d["clocks"])
if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"])
