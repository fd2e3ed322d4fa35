for v in "L1 RETTO_B200_LANES=1" "L2u256 RETTO_B200_LANES=2 RETTO_B200_UNIT_PAGES=256" "L2u128 RETTO_B200_LANES=2 RETTO_B200_UNIT_PAGES=128"; do
  set -- $v
  tag=$1; shift
  env $@ timeout 400 python bench.py --no-cpu-baseline --pages 768 > gpurun_out/y_$tag.json 2> gpurun_out/y_$tag.err
  python tools/show_bench.py gpurun_out/y_$tag.json > gpurun_out/y_$tag.txt 2>&1
  head -1 gpurun_out/y_$tag.txt
done
