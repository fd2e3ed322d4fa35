import os, sys, subprocess, json
for env in ({}, {"RETTO_B200_NO_NF_PROBE": "1"}):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-forward", "--no-variants", "--steps", "10"], capture_output=True, text=True, env=e)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    print(env, "value", round(d["value"]), {k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items() if k.startswith("bitmap") or k.startswith("build") or k.startswith("crop_rows")})
