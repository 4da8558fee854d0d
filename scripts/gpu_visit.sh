# One GPU-box visit: parity tests (incl. the config[1]-scale case, optionally at the full 1 M pairs), the bench line,
# the ncu launch list and `--set full` captures of $KERNELS exported as raw CSV + CUDA/SASS source CSV.
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
if [ -n "$FULL_PAIRS" ]; then
( time B200BWA_FULL_PAIRS=$FULL_PAIRS python -m pytest tests -m gpu -x -q -k config1_scale ) > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -5 gpurun_out/pytest_full.log
fi
( time python bench.py --steps 3 --warmup 3 ${BENCH_ARGS} ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
if [ -z "$SKIP_REF" ]; then
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
if [ -z "$SKIP_LAUNCHES" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
fi
for k in ${KERNELS}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r1_${k} python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /tmp/ncu_${k}.log 2>&1
tail -5 /tmp/ncu_${k}.log > gpurun_out/ncu_${k}.log
ncu -i /tmp/r1_${k}.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>/dev/null
ncu -i /tmp/r1_${k}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${k}_source.csv.gz
ncu -i /tmp/r1_${k}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > gpurun_out/${k}_cuda_sass.csv.gz
done
du -sh gpurun_out
