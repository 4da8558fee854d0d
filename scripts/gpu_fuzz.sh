#!/bin/bash
# Differential fuzzing of the PRODUCT binary (b200bwa -> libbwa.so on cuda:0) against the oracle, time-boxed.
O=gpurun_out/${OUTDIR:-fuzz}
mkdir -p $O
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader > $O/gpu.txt
timeout 200 python scripts/fuzz_parity.py --engine gpu --first 20000 --count 100000 --jobs ${JOBS:-8} --seconds ${T1:-100} --out /tmp/fz1 > $O/fuzz_regular.log 2>&1
tail -1 $O/fuzz_regular.log
timeout 200 python scripts/fuzz_parity.py --engine gpu --first 2000000 --count 100000 --jobs ${JOBS:-8} --seconds ${T2:-100} --out /tmp/fz2 > $O/fuzz_extreme.log 2>&1
tail -1 $O/fuzz_extreme.log
cp /tmp/fz1/report.jsonl $O/report_regular.jsonl; cp /tmp/fz2/report.jsonl $O/report_extreme.jsonl
# keep the inputs of any finding (small)
for d in /tmp/fz1/t* /tmp/fz2/t*; do [ -d "$d" ] && tar czf $O/$(basename $d).tgz -C $(dirname $d) $(basename $d) 2>/dev/null; done
ls $O | head -30
