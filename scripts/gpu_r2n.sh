#!/bin/bash
# Round-2 visit N (1 GPU): parity of HEAD + bench + launch-configuration sweep through the engine's environment knobs.
O=gpurun_out/${OUTDIR:-r2n}
mkdir -p $O
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
fi
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print(d['stage_ms_per_step'])"
: > $O/sweep.jsonl
run() { name=$1; shift; echo "{\"variant\": \"$name\"}" >> $O/sweep.jsonl; env "$@" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu $BARGS >> $O/sweep.jsonl 2> $O/sweep_$name.err; BARGS=; }
run base A=1
run base2 A=1
run seedblk64 B200BWA_SEED_BLOCK=64 B200BWA_SEED_BPS=8 B200BWA_SEED3_BPS=16
run seedblk32 B200BWA_SEED_BLOCK=32 B200BWA_SEED_BPS=16 B200BWA_SEED3_BPS=32
run seedgrid296 B200BWA_SEED_GRID=296
run seedgrid444 B200BWA_SEED_GRID=444
run laneblk64 B200BWA_LANE_BLOCK=64 B200BWA_LANE_GRID=1184
run stagger2 B200BWA_BENCH_STAGGER_MS=2
run stagger8 B200BWA_BENCH_STAGGER_MS=8
BARGS="--workers 8"; run w8stagger4 B200BWA_WORKERS=8 B200BWA_BENCH_STAGGER_MS=4
python - <<P
import json
name=None
for l in open('$O/sweep.jsonl'):
    d=json.loads(l)
    if 'variant' in d: name=d['variant']; continue
    print(name, round(d['value']/1e6,3), 'ms/step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})
P
