#!/bin/bash
# Round-2 visit P (1 GPU): heavy/light split of the pair finalisation -- parity, bench, threshold sweep.
O=gpurun_out/${OUTDIR:-r2p}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print(d['stage_ms_per_step'])"
grep "last batch delivered\|threads joined\|unmapped\|truncated" $O/bench_n1.err | tail -8
: > $O/sweep.jsonl
run() { name=$1; shift; echo "{\"variant\": \"$name\"}" >> $O/sweep.jsonl; env "$@" python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu $BARGS >> $O/sweep.jsonl 2> $O/sweep_$name.err; BARGS=; }
for h in ${HEAVY:-0 3 4 6 10 20 6 0}; do run heavy$h B200BWA_FINAL_HEAVY=$h; done
python - <<P
import json
name=None
for l in open('$O/sweep.jsonl'):
    d=json.loads(l)
    if 'variant' in d: name=d['variant']; continue
    print(name, round(d['value']/1e6,3), 'ms/step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})
P
B200BWA_DEBUG_TIMES=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000 --workers 1 > /dev/null 2> $O/final_times.err; grep "D::final" $O/final_times.err | head -60 > $O/final_times.txt; rm -f $O/final_times.err
