#!/bin/bash
# Last GPU visit of the round (1 GPU, ~3.5 min): (1) the product binary against the oracle with deliberately tiny
# per-lane regions / pool slots / interval and SAM-slot capacities, so that random inputs drive the device-side overflow
# and retry paths (where round 2's only device fault lived); (2) the bench command without a profiler, then its ncu
# launch list (the build with k_pair_names).
O=gpurun_out/${OUTDIR:-final4}
mkdir -p $O
timeout 150 python -m pytest tests/test_gpu.py -q -k "capacity_retry_paths_on_device or paired_names" > $O/pytest_new.log 2>&1; echo "pytest(new tests) rc $?"; tail -3 $O/pytest_new.log
FUZZ_ENGINE_ENV="B200BWA_SCRATCH=1024,B200BWA_BIG_SLOT=4096,B200BWA_BIG_POOL=1048576,B200BWA_CAP_INTV=2,B200BWA_SLOT=64" \
  timeout 170 python scripts/fuzz_parity.py --engine gpu --first 1300000 --count 100000 --jobs 8 --seconds ${T1:-80} --out /tmp/fz3 > $O/fuzz_tiny_capacities.log 2>&1
tail -1 $O/fuzz_tiny_capacities.log; cp /tmp/fz3/report.jsonl $O/report_tiny_capacities.jsonl
for d in /tmp/fz3/t*; do [ -d "$d" ] && tar czf $O/$(basename $d).tgz -C /tmp/fz3 $(basename $d) 2>/dev/null; done
SMALL="--steps 1 --warmup 1 --no-e2e --no-cpu --pairs 400000"
python bench.py $SMALL > $O/bench_small.json 2> $O/bench_small.err; echo "plain rc $?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 900 --csv --log-file $O/launches.csv python bench.py $SMALL > $O/ncu_launch.log 2>&1; echo "launch list rc $?"
ls -la $O | head -20
