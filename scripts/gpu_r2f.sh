#!/bin/bash
# Round-2 visit F (1 GPU): configs[2] at its stated reference size -- 3.1 Gbp index built on the GPU, 20 M pairs.
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
python scripts/cfg2_run.py --ref-len ${1:-3100000000} --pairs ${2:-20000000} --prefix-batches 12 --late-batches 2 --out $O > $O/cfg2.log 2> $O/cfg2.err; echo "cfg2 rc $?"; tail -40 $O/cfg2.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $O/mem.txt
