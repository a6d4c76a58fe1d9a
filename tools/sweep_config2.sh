#!/bin/bash
# BASELINE config 2: batch-size sweep 1K-64K x bilinear_type on one B200 (one JSON line per point).
#   bash tools/sweep_config2.sh > profiles/r1_sweep_config2.jsonl
set -u
cd "$(dirname "$0")/.."
for B in 1024 2048 4096 8192 16384 32768 65536; do
  python bench.py --batch $B --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{'
done
for T in each interaction; do
  for B in 4096 16384 65536; do
    python bench.py --batch $B --bilinear $T --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{'
  done
done
python bench.py --batch 16384 --precision bf16 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{'
python bench.py --batch 65536 --precision bf16 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{'
python bench.py --batch 16384 --id-dist zipf --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{'
python bench.py --mode infer --batch 8192 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{'
python bench.py --mode infer --batch 65536 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{'
