#!/bin/bash
# Round-2 visit A (1 GPU): parity tests, INT/DPX peaks, bench (engine + reference arm), launch list.
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi -L > $O/gpus.txt 2>&1; nproc >> $O/gpus.txt
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
scripts/_build/int_peak > $O/int_peaks.json 2> $O/int_peaks.err; cat $O/int_peaks.json
B200BWA_TIMES=1 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; cat $O/bench_n1.json
grep "T::b200bwa" $O/bench_n1.err | tail -80 > $O/bench_n1_times.txt
python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --pairs 200000 > $O/ncu.log 2>&1; echo "ncu rc $?"
