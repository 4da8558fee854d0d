#!/bin/bash
# Round-2 visit E (1 GPU): full GPU test suite incl. the 100 Mbp index parity, bench (mmap output), configs[2]-class run.
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
df -h /dev/shm > $O/env.txt; free -g >> $O/env.txt; nproc >> $O/env.txt
B200BWA_FULL_INDEX=1 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -14 $O/pytest_gpu.log
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'], 'cpu', d.get('cpu_baseline', {}).get('value'))"
grep "T::b200bwa" $O/bench_n1.err | tail -150 > $O/bench_n1_times.txt
python scripts/cfg2_run.py --ref-len ${1:-1000000000} --pairs ${2:-8600000} --out $O > $O/cfg2.log 2> $O/cfg2.err; echo "cfg2 rc $?"; tail -25 $O/cfg2.err
