#!/bin/bash
# Round-2 visit U (1 GPU): in-place digits + 4-positions-per-lane MD loop; 6 vs 8 stream sets end to end; seeding step
# distribution; launch list + --set full captures of this build for profiles/.
O=gpurun_out/${OUTDIR:-r2u}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'], 'cpu', d.get('cpu_baseline')); print(d['stage_ms_per_step'])"
B200BWA_WORKERS=8 python bench.py --steps 3 --warmup 3 --no-cpu --workers 8 > $O/bench_n1_w8.json 2> $O/bench_n1_w8.err; python -c "
import json; d=json.load(open('$O/bench_n1_w8.json')); print('w8 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'])"
B200BWA_WORKERS=7 python bench.py --steps 3 --warmup 3 --no-cpu --workers 7 > $O/bench_n1_w7.json 2> $O/bench_n1_w7.err; python -c "
import json; d=json.load(open('$O/bench_n1_w7.json')); print('w7 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_call'])"
python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc $?"; cut -c1-300 $O/bench_ref.json
SMALL="--steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000"
B200BWA_DEBUG_TIMES=1 python bench.py $SMALL --workers 1 > /dev/null 2> $O/dbg.err; grep "D::seed\|D::final" $O/dbg.err | head -30 > $O/dbg_times.txt; rm -f $O/dbg.err; grep "D::seed" $O/dbg_times.txt | head -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 900 --csv --log-file $O/launches.csv python bench.py $SMALL > $O/ncu_launch.log 2>&1; echo "launch list rc $?"
for k in ${KERNELS:-k_seed k_final_pe_heavy k_final_pe k_xext_run k_chain}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r2_${k} python bench.py $SMALL > /tmp/ncu_${k}.log 2>&1
  ncu -i /tmp/r2_${k}.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_${k}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $O/${k}_cuda_sass.csv.gz
done
