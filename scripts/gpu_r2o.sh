#!/bin/bash
# Round-2 visit O (1 GPU): fresh ncu evidence of the current code (launch list + --set full of the stage kernels) and the
# per-pair cycle distribution of k_final_pe (B200BWA_DEBUG_TIMES).  Every profiled command has run without ncu first.
O=gpurun_out/${OUTDIR:-r2o}
mkdir -p $O
SMALL="--steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000"
python bench.py $SMALL > $O/bench_small.json 2> $O/bench_small.err; echo "bench rc $?"
B200BWA_DEBUG_TIMES=1 python bench.py $SMALL --workers 1 > /dev/null 2> $O/final_times.err; grep "D::final" $O/final_times.err | head -400 > $O/final_times.txt; rm -f $O/final_times.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 700 --csv --log-file $O/launches.csv python bench.py $SMALL > $O/ncu_launch.log 2>&1; echo "launch list rc $?"
for k in ${KERNELS:-k_final_pe k_seed k_chain k_extend k_ext_chain k_galn k_xext_run k_resc_sw k_seed3}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r2_${k} python bench.py $SMALL > /tmp/ncu_${k}.log 2>&1
  tail -3 /tmp/ncu_${k}.log > $O/ncu_${k}.log
  ncu -i /tmp/r2_${k}.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_${k}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $O/${k}_cuda_sass.csv.gz
done
du -sh $O
