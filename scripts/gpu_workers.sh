# Sweep of the number of batches in flight (worker threads / streams): one bench line per setting.
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/workers.jsonl
for W in ${WORKERS:-4 6 8}; do
  B200BWA_WORKERS=$W python bench.py --workers $W --steps 3 --warmup 3 --no-cpu >> gpurun_out/workers.jsonl 2> gpurun_out/workers_$W.err
done
cut -c1-300 gpurun_out/workers.jsonl
