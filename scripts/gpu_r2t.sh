#!/bin/bash
# Round-2 visit T (1 GPU): pipelined k_sa, pairs needing an in-kernel DP routed to k_final_pe_heavy, 16 KB scratch regions.
O=gpurun_out/${OUTDIR:-r2t}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 B200BWA_BENCH_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print(d['stage_ms_per_step']); print(d['roofline']['sa_kernel'])"
grep "bench trace" $O/bench_n1.err | grep "total\|finalize"
: > $O/sweep.jsonl
run() { name=$1; shift; echo "{\"variant\": \"$name\"}" >> $O/sweep.jsonl; env "$@" B200BWA_BENCH_TRACE=1 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu $BARGS >> $O/sweep.jsonl 2> $O/sweep_$name.err; grep "bench trace" $O/sweep_$name.err > $O/trace_$name.txt; BARGS=; }
run default A=1
run nosaplan B200BWA_NO_SA_PLAN=1
run default2 A=1
BARGS="--workers 8"; run w8 B200BWA_WORKERS=8
python - <<P
import json
name=None
for l in open('$O/sweep.jsonl'):
    d=json.loads(l)
    if 'variant' in d: name=d['variant']; continue
    print(name, round(d['value']/1e6,3), 'ms/step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})
P
SMALL="--steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000"
B200BWA_DEBUG_TIMES=1 python bench.py $SMALL --workers 1 > /dev/null 2> $O/final_times.err; grep "D::final" $O/final_times.err | head -24 > $O/final_times.txt; rm -f $O/final_times.err; head -3 $O/final_times.txt
for k in ${KERNELS:-k_sa k_final_pe}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r2_${k} python bench.py $SMALL > /tmp/ncu_${k}.log 2>&1
  ncu -i /tmp/r2_${k}.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_${k}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $O/${k}_cuda_sass.csv.gz
done
