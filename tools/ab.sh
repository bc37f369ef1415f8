#!/bin/bash
# A/B timing of tuning builds (wobble_jax_b200/alt_*.so) on a batch of chunks; run on the GPU box.
# usage: tools/ab.sh [n_chunks]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=${1:-16}
{
for alt in wobble_jax_b200/alt_*.so; do
  echo "== $alt"; JABBLE_B200_LIB=$PWD/$alt python tools/batch_check.py $n | tail -1
done
} 2>&1 | tee -a gpurun_out/ab.log
