timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
for v in "A"; do
  set -- $v
  tag=$1; shift
  env $@ timeout 300 python bench.py --no-cpu-baseline > gpurun_out/x_$tag.json 2> gpurun_out/x_$tag.err
  python tools/show_bench.py gpurun_out/x_$tag.json > gpurun_out/x_$tag.txt 2>&1
  cat gpurun_out/x_$tag.txt
done
