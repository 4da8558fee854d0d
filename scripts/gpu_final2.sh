#!/bin/bash
# Single-GPU check of the round's last build: pytest -m gpu, smoke(), default bench, then the time-boxed GPU fuzz.
O=gpurun_out/${OUTDIR:-final2}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?"; tail -1 $O/smoke.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'], 'launches', d['gpu_launches']); print(d['roofline']['frac'], d['clocks'])"
OUTDIR=${OUTDIR:-final2}/fuzz T1=${T1:-70} T2=${T2:-70} bash scripts/gpu_fuzz.sh
