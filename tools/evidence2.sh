#!/bin/bash
# second evidence pass on ONE B200: ncu capture of the decode kernels, side configs with / without the ccl tiers, mixed workload, sanitizer
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jpeg.py tests/test_gpu_det_post.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none -k regex:"jpeg_" --launch-skip 12 --launch-count 4 -f -o gpurun_out/jpeg_prof python tools/bench_jpeg.py 256 1280 > gpurun_out/jpeg_prof.log 2>&1
python tools/ncu_traffic.py gpurun_out/jpeg_prof.ncu-rep gpurun_out/final_traffic_jpeg.json "ncu --set full --clock-control none, tools/bench_jpeg.py 256 1280, the 4 decode kernels of the 4th decode call (restart interval per MCU row)" gpurun_out/final_ncu_full_raw_jpeg.csv
rm -f gpurun_out/jpeg_prof.ncu-rep
python tools/bench_configs.py 2>/dev/null > gpurun_out/final_configs.json; head -1 gpurun_out/final_configs.json | cut -c1-900
RETTO_B200_CCL_ONE_TIER=1 python tools/bench_configs.py 2>/dev/null | head -1 | cut -c1-900
timeout 600 python bench.py --workload mixed --pages 2048 --unique 64 --no-cpu-baseline > gpurun_out/final_mixed.json 2> gpurun_out/final_mixed.err; python tools/show_bench.py gpurun_out/final_mixed.json 2>&1 | head -14
timeout 300 python tools/bench_jpeg.py 256 1280 > gpurun_out/final_jpeg.txt 2>&1; cp gpurun_out/r02_bench_jpeg.json gpurun_out/final_jpeg.json; cut -c1-400 gpurun_out/final_jpeg.txt
bash tools/sanitize.sh
