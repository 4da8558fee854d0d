# Sweep of one environment knob: VAR=name VALUES="a b c" -> one bench line per value in gpurun_out/envsweep.jsonl
set -x
mkdir -p gpurun_out
: > gpurun_out/envsweep.jsonl
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
for v in $VALUES; do
  echo "{\"var\": \"$VAR\", \"value_set\": \"$v\"}" >> gpurun_out/envsweep.jsonl
  env $VAR=$v python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu >> gpurun_out/envsweep.jsonl 2> gpurun_out/envsweep_$v.err
done
