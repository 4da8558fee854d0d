#!/bin/bash
# Round-2 visit Q (1 GPU): shared-memory arena + unrolled bulk copies in the finalisation, one-fetch LF step in k_sa,
# query-word runs in k_xext_run -- parity, bench, A/B of the switches, per-batch trace of the concurrent passes.
O=gpurun_out/${OUTDIR:-r2q}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print(d['stage_ms_per_step']); print(json.dumps(d['sw']['kernels'])); print(d['roofline'])"
: > $O/sweep.jsonl
run() { name=$1; shift; echo "{\"variant\": \"$name\"}" >> $O/sweep.jsonl; env "$@" B200BWA_BENCH_TRACE=1 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu $BARGS >> $O/sweep.jsonl 2> $O/sweep_$name.err; grep "bench trace" $O/sweep_$name.err > $O/trace_$name.txt; BARGS=; }
run default A=1
run fast0 B200BWA_FINAL_FAST=0
run heavy0 B200BWA_FINAL_HEAVY=0
run heavy3 B200BWA_FINAL_HEAVY=3
run heavy10 B200BWA_FINAL_HEAVY=10
run default2 A=1
python - <<P
import json
name=None
for l in open('$O/sweep.jsonl'):
    d=json.loads(l)
    if 'variant' in d: name=d['variant']; continue
    print(name, round(d['value']/1e6,3), 'ms/step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})
P
for v in default heavy10; do echo "== trace $v"; cat $O/trace_$v.txt; done
B200BWA_DEBUG_TIMES=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000 --workers 1 > /dev/null 2> $O/final_times.err; grep "D::final" $O/final_times.err | head -40 > $O/final_times.txt; rm -f $O/final_times.err; head -3 $O/final_times.txt
