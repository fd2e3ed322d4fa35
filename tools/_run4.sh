timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
RETTO_B200_BR3_NW=8 timeout 900 python -m pytest tests/test_gpu_det_post.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/gputests8.log 2>&1; tail -3 gpurun_out/gputests8.log
for v in "A" "B RETTO_B200_BR3_NW=8" "C RETTO_B200_BR2=1"; do
  set -- $v
  tag=$1; shift
  env $@ timeout 300 python bench.py --no-cpu-baseline > gpurun_out/x_$tag.json 2> gpurun_out/x_$tag.err
  python tools/show_bench.py gpurun_out/x_$tag.json > gpurun_out/x_$tag.txt 2>&1
  head -1 gpurun_out/x_$tag.txt; grep -E "bitmap|collapse|kernel sum" gpurun_out/x_$tag.txt
done
