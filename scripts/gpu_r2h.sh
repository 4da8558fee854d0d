#!/bin/bash
# Round-2 visit H (1 GPU): ncu evidence -- launch list of the bench command, `--set full` captures of the stage kernels at
# configs[1] (100 Mbp) and of k_seed / k_sa at a 1 Gbp index (first seeding-vs-HBM numbers beyond the L2), raw CSV +
# CUDA/SASS source pages.  Every profiled command has run once without ncu first (bench line kept next to the captures).
O=gpurun_out/r2h
mkdir -p $O
python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 700 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000 > $O/ncu_launch.log 2>&1; echo "launch list rc $?"
for k in ${KERNELS:-k_seed k_sa k_final_pe k_resc_sw k_xext_run k_chain k_ingest_copy}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r2_${k} python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000 > /tmp/ncu_${k}.log 2>&1
  tail -3 /tmp/ncu_${k}.log > $O/ncu_${k}.log
  ncu -i /tmp/r2_${k}.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_${k}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $O/${k}_cuda_sass.csv.gz
done
# ---- 1 Gbp index: seeding and SA lookup against HBM ----
BIG="--ref-len 1000000000 --pairs 300000 --steps 2 --warmup 2 --no-cpu"
python bench.py $BIG > $O/bench_1Gbp.json 2> $O/bench_1Gbp.err; echo "bench 1Gbp rc $?"; tail -2 $O/bench_1Gbp.err
for k in k_seed k_sa; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r2_${k}_1G python bench.py $BIG --steps 1 --warmup 1 --no-e2e > /tmp/ncu_${k}_1G.log 2>&1
  tail -3 /tmp/ncu_${k}_1G.log > $O/ncu_${k}_1Gbp.log
  ncu -i /tmp/r2_${k}_1G.ncu-rep --page raw --csv > $O/${k}_1Gbp_raw.csv 2>/dev/null
done
# index builder kernels at 1 Gbp (one radix pass pair)
python - <<'EOF' > gpurun_out/r2h/index_1Gbp_fa.log 2>&1
import sys, time
sys.path.insert(0, ".")
from sparkbwa_b200 import synth
ctg = synth.make_reference(1_000_000_000, n_contigs=8, seed=1, repeat_frac=0.02)
synth.write_fasta("/dev/shm/ix1g.fa", ctg)
EOF
timeout 900 ncu --set full --clock-control none -k regex:^k_ix_rs_ -s 8 -c 2 -f -o /tmp/r2_ix sparkbwa_b200/b200bwa index -p /dev/shm/ix1g /dev/shm/ix1g.fa > $O/ncu_index.log 2>&1
ncu -i /tmp/r2_ix.ncu-rep --page raw --csv > $O/k_ix_rs_1Gbp_raw.csv 2>/dev/null
rm -f /dev/shm/ix1g*
du -sh $O
