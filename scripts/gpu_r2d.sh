#!/bin/bash
# Round-2 visit D (N GPUs, default 2): one job over N GPUs -- parity test on all devices + strong-scaling bench.
N=${1:-2}
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
nvidia-smi -L > $O/gpus_n$N.txt; nproc >> $O/gpus_n$N.txt
python -m pytest tests/test_gpu.py -m gpu -x -q -k "several_engines or config1 or repeat_call" > $O/pytest_n$N.log 2>&1; echo "pytest rc $?" >> $O/pytest_n$N.log
tail -3 $O/pytest_n$N.log
B200BWA_TIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('$O/bench_n$N.json')); print('N=$N value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'])"
grep "T::b200bwa" $O/bench_n$N.err | tail -120 > $O/bench_n${N}_times.txt
