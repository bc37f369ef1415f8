#!/bin/bash
# A/B timing of the fused kernel variants on one chunk (tools/prof_one.py); run on the GPU box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== v1 (general kernel)"; JB_FUSED_V1=1 python tools/prof_one.py --evals 8
echo "== v2 r128"; python tools/prof_one.py --evals 8
for alt in wobble_jax_b200/alt_*.so; do
  echo "== v2 $alt"; JABBLE_B200_LIB=$PWD/$alt python tools/prof_one.py --evals 8
done
} 2>&1 | tee gpurun_out/ab.log
