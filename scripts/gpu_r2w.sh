#!/bin/bash
# Round-2 visit W (1 GPU): lane compaction in k_seed -- parity, bench, span / grid sweep, ncu of k_seed.
O=gpurun_out/${OUTDIR:-r2w}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print(d['stage_ms_per_step']); print(d['roofline'])"
: > $O/sweep.jsonl
run() { name=$1; shift; echo "{\"variant\": \"$name\"}" >> $O/sweep.jsonl; env "$@" python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu $BARGS >> $O/sweep.jsonl 2> $O/sweep_$name.err; BARGS=; }
run span128 A=1
run span0 B200BWA_SEED_SPAN=0
run span64 B200BWA_SEED_SPAN=64
run span256 B200BWA_SEED_SPAN=256
run span128grid444 B200BWA_SEED_GRID=444
run span64grid444 B200BWA_SEED_GRID=444 B200BWA_SEED_SPAN=64
run span128heavy8 B200BWA_FINAL_HEAVY=8
python - <<P
import json
name=None
for l in open('$O/sweep.jsonl'):
    d=json.loads(l)
    if 'variant' in d: name=d['variant']; continue
    print(name, round(d['value']/1e6,3), 'ms/step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items() if k in ('seed','finalize','total')})
P
SMALL="--steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000"
for k in k_seed; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r2_${k} python bench.py $SMALL > /tmp/ncu_${k}.log 2>&1
  ncu -i /tmp/r2_${k}.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_${k}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $O/${k}_cuda_sass.csv.gz
done
python scripts/ncu_summary.py $O/k_seed_raw.csv | grep "time_duration\|inst_executed.sum\|thread_inst_executed_per\|issue_active\|cycles_active.avg\|cycles_elapsed.max\|warps_active\|registers"
