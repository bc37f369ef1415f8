#!/bin/bash
# A/B timing of the fused kernel variants on one chunk (tools/prof_one.py); run on the GPU box.
# usage: tools/ab.sh [extra prof_one args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for alt in wobble_jax_b200/alt_*.so; do
  echo "== $alt $*"; JABBLE_B200_LIB=$PWD/$alt python tools/prof_one.py --evals 8 "$@"
done
} 2>&1 | tee -a gpurun_out/ab.log
