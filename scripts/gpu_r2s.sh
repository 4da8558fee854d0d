#!/bin/bash
# Round-2 visit S (1 GPU): outlier pool (tasks that overflow their scratch region re-run in a pool slot) -- parity, bench,
# per-lane region size sweep, per-batch trace (allocation stalls).
O=gpurun_out/${OUTDIR:-r2s}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 B200BWA_BENCH_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print(d['stage_ms_per_step'])"
grep "bench trace" $O/bench_n1.err
: > $O/sweep.jsonl
run() { name=$1; shift; echo "{\"variant\": \"$name\"}" >> $O/sweep.jsonl; env "$@" B200BWA_BENCH_TRACE=1 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu $BARGS >> $O/sweep.jsonl 2> $O/sweep_$name.err; grep "bench trace" $O/sweep_$name.err > $O/trace_$name.txt; BARGS=; }
run default A=1
run s16k B200BWA_SCRATCH=16384
run s24k B200BWA_SCRATCH=24576
run s8k B200BWA_SCRATCH=8192
run nopool B200BWA_BIG_POOL=0
run default2 A=1
python - <<P
import json
name=None
for l in open('$O/sweep.jsonl'):
    d=json.loads(l)
    if 'variant' in d: name=d['variant']; continue
    print(name, round(d['value']/1e6,3), 'ms/step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})
P
for v in default s16k s8k nopool; do echo "== trace $v"; grep "total\|chain\|finalize" $O/trace_$v.txt; done
nvidia-smi --query-gpu=memory.used --format=csv
