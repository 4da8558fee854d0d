# Quick GPU visit: parity tests + one bench line (+ optional ncu of $KERNELS).
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time python bench.py --steps 3 --warmup 3 ${BENCH_ARGS} ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
for k in ${KERNELS}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r1_${k} python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /tmp/ncu_${k}.log 2>&1
tail -5 /tmp/ncu_${k}.log > gpurun_out/ncu_${k}.log
ncu -i /tmp/r1_${k}.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>/dev/null
ncu -i /tmp/r1_${k}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${k}_source.csv.gz
done
du -sh gpurun_out
