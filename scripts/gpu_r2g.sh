#!/bin/bash
# Round-2 visit G (N GPUs, default 8): strong-scaling bench at N + configs[2] (3.1 Gbp, 20 M pairs) as ONE job over N GPUs.
N=${1:-8}
mkdir -p gpurun_out/r2g
O=gpurun_out/${OUTDIR:-r2g}
nvidia-smi -L > $O/gpus_n$N.txt; nproc >> $O/gpus_n$N.txt; free -g >> $O/gpus_n$N.txt
B200BWA_TIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('$O/bench_n$N.json')); print('N=$N value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'], 'large', d.get('e2e_large_job'))"
grep "T::b200bwa" $O/bench_n$N.err | tail -150 > $O/bench_n${N}_times.txt
python scripts/cfg2_run.py --ref-len ${2:-3100000000} --pairs ${3:-20000000} --prefix-batches 12 --late-batches 2 --out $O > $O/cfg2.log 2> $O/cfg2.err; echo "cfg2 rc $?"; grep "cfg2 +" $O/cfg2.err
