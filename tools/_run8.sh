RETTO_B200_CROP_VEC=1 RETTO_B200_BR3_MINB=10 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
for v in "A" "B RETTO_B200_CROP_VEC=1" "C RETTO_B200_BR3_MINB=10" "D RETTO_B200_BR3_MINB=12" "E RETTO_B200_CROP_VEC=1"  "F"; do
  set -- $v
  tag=$1; shift
  env $@ timeout 300 python bench.py --no-cpu-baseline > gpurun_out/x_$tag.json 2> gpurun_out/x_$tag.err
  python tools/show_bench.py gpurun_out/x_$tag.json > gpurun_out/x_$tag.txt 2>&1
  echo "== $tag $@"; head -1 gpurun_out/x_$tag.txt; grep -E "crop_rows|bitmap|kernel sum" gpurun_out/x_$tag.txt
done
