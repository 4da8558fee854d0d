#!/bin/bash
# Round-2 visit I (1 GPU): parity after the rescue-SW / scan / output-path changes + bench.
O=gpurun_out/${OUTDIR:-r2i}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print(json.dumps(d['sw']['kernels'])); print(d['stage_ms_per_step'])"
