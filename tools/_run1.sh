set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for v in "A" "B RETTO_B200_BR2_DEPTH=2" "C RETTO_B200_BR2_DEPTH=3" "D RETTO_B200_BR2_DEPTH=4" "E RETTO_B200_BB_DIRECT=1" "F RETTO_B200_GEOM_SERIAL=1"; do
  set -- $v
  tag=$1; shift
  env $@ timeout 300 python bench.py --no-cpu-baseline > gpurun_out/x_$tag.json 2> gpurun_out/x_$tag.err
  python tools/show_bench.py gpurun_out/x_$tag.json | head -12
done
