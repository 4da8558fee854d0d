#!/bin/bash
# Round-2 visit C (1 GPU): parity incl. forced in-kernel DP paths, bench with fixed scratch policy + async writers.
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'], 'cpu', d.get('cpu_baseline', {}).get('value'))"
grep "T::b200bwa" $O/bench_n1.err | tail -150 > $O/bench_n1_times.txt
B200BWA_WORKERS=8 python bench.py --steps 3 --warmup 3 --no-cpu --workers 8 > $O/bench_n1_w8.json 2> $O/bench_n1_w8.err; python -c "
import json; d=json.load(open('$O/bench_n1_w8.json')); print('workers 8 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'])"
B200BWA_BLOCKING_SYNC=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1_block.json 2> $O/bench_n1_block.err; python -c "
import json; d=json.load(open('$O/bench_n1_block.json')); print('blocking sync value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'])"
