#!/bin/bash
# compute-sanitizer over the GPU suite (everything but the full-size tests) on ONE B200: logs -> gpurun_out/r02_sanitizer_*.log
what=${1:-all}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
if [ "$what" = all ] || [ "$what" = mem ]; then
timeout 1500 $S --tool memcheck --leak-check no --error-exitcode 9 python -m pytest tests/test_gpu_golden.py tests/test_gpu_stage1.py tests/test_gpu_det_post.py tests/test_gpu_pipeline.py tests/test_gpu_jpeg.py tests/test_gpu_configs.py -q -m gpu -x -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -4 gpurun_out/r02_sanitizer_memcheck.log
fi
if [ "$what" = all ] || [ "$what" = race ]; then
timeout 900 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py "tests/test_gpu_det_post.py::test_hole_borders" "tests/test_gpu_jpeg.py::test_decode_matrix" "tests/test_gpu_pipeline.py::test_batches_and_cls_flip_parity" -q -m gpu -x -p no:cacheprovider > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -4 gpurun_out/r02_sanitizer_racecheck.log
fi
