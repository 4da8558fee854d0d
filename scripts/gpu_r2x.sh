#!/bin/bash
# Round-2 visit X (N GPUs, default 2): multi-engine parity tests + strong-scaling bench at N.
N=${1:-2}
O=gpurun_out/${OUTDIR:-r2x}
mkdir -p $O
nvidia-smi -L > $O/gpus_n$N.txt; nproc >> $O/gpus_n$N.txt
python -m pytest tests -m gpu -x -q -k "several_engines or concurrent or partitions or cabi" > $O/pytest_n$N.log 2>&1; echo "pytest rc $?" >> $O/pytest_n$N.log; tail -2 $O/pytest_n$N.log
B200BWA_TIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('$O/bench_n$N.json')); print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call']); print('large', d.get('e2e_large_job')); print('partitions', d.get('e2e_partitions'))"
grep "T::b200bwa" $O/bench_n$N.err | grep -v "worker\|cut at\|queued" | tail -12 > $O/bench_n${N}_times.txt
