#!/bin/bash
# round-end evidence run on ONE B200 (writes gpurun_out/final_*).  bash tools/final_profile.sh [bench|ncu|all]
# gpurun copies back at most 64 MiB: the full ncu capture is exported to CSV / JSON on the box and the .ncu-rep dropped when large.
what=${1:-all}
mkdir -p gpurun_out
if [ "$what" = bench ] || [ "$what" = all ]; then
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_gputests.log 2>&1; tail -2 gpurun_out/final_gputests.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python tools/show_bench.py gpurun_out/final_bench.json > gpurun_out/final_bench.txt 2>&1; head -1 gpurun_out/final_bench.txt
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 300 python tools/bench_configs.py > gpurun_out/final_configs.json 2> gpurun_out/final_configs.err
timeout 600 python bench.py --workload mixed --pages 2048 --unique 64 --no-cpu-baseline > gpurun_out/final_mixed.json 2> gpurun_out/final_mixed.err
timeout 300 python tools/bench_jpeg.py 256 1280 > gpurun_out/final_jpeg.txt 2>&1; cp gpurun_out/r02_bench_jpeg.json gpurun_out/final_jpeg.json
fi
if [ "$what" = ncu ] || [ "$what" = all ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --no-forward > gpurun_out/final_bench_under_ncu.log 2>&1
# full capture: the e2e (JPEG) pass runs every kernel of the path once per step; skip the set-up and the value pass's warm-up
timeout 1200 ncu --set full --clock-control none -k regex:"bitmap_runs3|ccl_runs|build_batches|crop_rows|crop_setup|box_geometry|box_score|ctc_argmax|ctc_collapse|det_pre_identity|page_sort|jpeg_" --launch-skip 120 --launch-count 40 -f -o gpurun_out/final_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --no-forward > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_ncu_full.log
python tools/ncu_traffic.py gpurun_out/final_prof.ncu-rep gpurun_out/final_traffic.json "ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --no-forward (256 pages 1280x1280), launches 120..159 of the kernel filter; raw page: final_ncu_full_raw.csv" gpurun_out/final_ncu_full_raw.csv
ls -la gpurun_out/final_prof.ncu-rep
if [ $(stat -c %s gpurun_out/final_prof.ncu-rep) -gt 40000000 ]; then rm gpurun_out/final_prof.ncu-rep; fi
fi
du -sh gpurun_out
