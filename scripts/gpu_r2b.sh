#!/bin/bash
# Round-2 visit B (1 GPU): parity tests incl. the new cases, INT/DPX peaks, bench e2e with async writers.
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
scripts/_build/int_peak > $O/int_peaks.json 2> $O/int_peaks.err; cat $O/int_peaks.json
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e'])"
grep "T::b200bwa" $O/bench_n1.err | tail -150 > $O/bench_n1_times.txt
for w in 8; do B200BWA_WORKERS=$w python bench.py --steps 3 --warmup 3 --no-cpu --workers $w > $O/bench_n1_w$w.json 2> $O/bench_n1_w$w.err; python -c "
import json; d=json.load(open('$O/bench_n1_w$w.json')); print('workers $w value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'])"; done
