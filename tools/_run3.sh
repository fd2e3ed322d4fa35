timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bitmap_runs2|ccl_runs|build_batches|crop_rows|box_geometry|ctc_argmax|det_pre_identity|ctc_collapse" --launch-skip 33 --launch-count 11 -f -o gpurun_out/prof_r01h python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_h.log 2>&1
tail -3 gpurun_out/ncu_full_h.log
