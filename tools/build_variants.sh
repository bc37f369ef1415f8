#!/bin/bash
# Tuning builds of libjabble_b200 (only the benchmark's k_fused2 variant): tools/build_variants.sh "<name>:<nvcc -D flags>" ...
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function --shared -cudart shared -DJB_F2_FAST_BUILD"
rm -f wobble_jax_b200/alt_*.so
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc $F $flags -Xptxas -v -o wobble_jax_b200/alt_$name.so wobble_jax_b200/csrc/jb_plan.cu > /tmp/build_$name.log 2>&1; grep -A2 "k_fused2ILi2ELi3ELi9" /tmp/build_$name.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $name" ) &
done
wait
