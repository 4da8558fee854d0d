#!/bin/bash
# ncu evidence of the round's last build (1 GPU): the bench command without a profiler first, then its launch list,
# then one --set full capture of k_seed (one launch = one 66 668-read batch of a 400 k-pair job).
O=gpurun_out/${OUTDIR:-final4}
mkdir -p $O
SMALL="--steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000"
python bench.py $SMALL > $O/bench_small.json 2> $O/bench_small.err; echo "plain rc $?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 900 --csv --log-file $O/launches.csv python bench.py $SMALL > $O/ncu_launch.log 2>&1; echo "launch list rc $?"
k=k_seed
timeout 120 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r2_${k} python bench.py $SMALL > /tmp/ncu_${k}.log 2>&1; echo "capture rc $?"
ncu -i /tmp/r2_${k}.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
ncu -i /tmp/r2_${k}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $O/${k}_cuda_sass.csv.gz
ls -la $O
