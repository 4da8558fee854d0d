# One GPU-box visit: parity tests, bench (engine + reference arm), ncu launch list, ncu --set full per kernel.
# ncu reports stay in /tmp on the box; only CSV exports (and the k_seed report) come back (gpurun_out <= 64 MiB).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi.txt
if [ -z "$SKIP_TESTS" ]; then
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
fi
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
if [ -z "$SKIP_REF" ]; then
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
for k in ${KERNELS:-k_seed k_final_pe k_ext_chain k_resc_sw k_extend k_chain k_sa}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 3 -c 1 -f -o /tmp/r1_${k} python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /tmp/ncu_${k}.log 2>&1
tail -5 /tmp/ncu_${k}.log > gpurun_out/ncu_${k}.log
ncu -i /tmp/r1_${k}.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>/dev/null
ncu -i /tmp/r1_${k}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${k}_source.csv.gz
done
cp /tmp/r1_k_seed.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
