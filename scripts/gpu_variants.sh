# Build-variant sweep on one GPU box: VARIANTS="name:lane_grid ..." -> one bench line per variant in gpurun_out/variants.jsonl
set -x
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1   # builds the index + reads once
for v in $VARIANTS; do
  n=${v%%:*}; g=${v#*:}
  lib=sparkbwa_b200/libbwa_$n.so
  [ "$n" = base ] && lib=sparkbwa_b200/libbwa.so
  echo "{\"variant\": \"$n\", \"lane_grid\": $g}" >> gpurun_out/variants.jsonl
  B200BWA_LIB=$PWD/$lib B200BWA_LANE_GRID=$g python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu >> gpurun_out/variants.jsonl 2> gpurun_out/variant_$n.err
done
cat gpurun_out/variants.jsonl | cut -c1-400
