#!/bin/bash
# Round-2 visit V (1 GPU): seeding launch shapes again, now that runs repeat within 1 % (no allocation stalls).
O=gpurun_out/${OUTDIR:-r2v}
mkdir -p $O
: > $O/sweep.jsonl
run() { name=$1; shift; echo "{\"variant\": \"$name\"}" >> $O/sweep.jsonl; env "$@" python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu $BARGS >> $O/sweep.jsonl 2> $O/sweep_$name.err; BARGS=; }
run base A=1
run seedblk32 B200BWA_SEED_BLOCK=32 B200BWA_SEED_BPS=16 B200BWA_SEED3_BPS=32
run seedblk64 B200BWA_SEED_BLOCK=64 B200BWA_SEED_BPS=8 B200BWA_SEED3_BPS=16
run seedgrid296 B200BWA_SEED_GRID=296
run seedgrid444 B200BWA_SEED_GRID=444
run blk32grid1480 B200BWA_SEED_BLOCK=32 B200BWA_SEED_BPS=16 B200BWA_SEED3_BPS=32 B200BWA_SEED_GRID=1480
run base2 A=1
python - <<P
import json
name=None
for l in open('$O/sweep.jsonl'):
    d=json.loads(l)
    if 'variant' in d: name=d['variant']; continue
    print(name, round(d['value']/1e6,3), 'ms/step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items() if k in ('seed','total')})
P
