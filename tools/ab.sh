#!/bin/bash
# A/B on ONE box: alternate bench.py between library variants (paths given as arguments), ROUNDS times each.
# usage: tools/ab.sh qwen3_asr_b200/libqasr_v0.so qwen3_asr_b200/libqasr_b200.so
set +e
mkdir -p gpurun_out
O=gpurun_out
ROUNDS=${ROUNDS:-2}
for r in $(seq 1 $ROUNDS); do
  i=0
  for lib in "$@"; do
    tag="ab_${i}_r${r}"
    QASR_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > $O/$tag.log 2> $O/$tag.err
    echo "== $lib round $r rc=$?"
    python tools/show_bench.py $O/$tag.log | grep -v "^clocks\|^cpu\|logmel\|proj[12]\|sum"
    i=$((i+1))
  done
done
