#!/bin/bash
# Timing ablations of the stacked fp32 leaf kernel (WRONG results by design): which stage of the pipeline paces it.
# HBSM_F32_MODE bits: 1 raw operand as hi (default), 4 no hi/lo split, 8 no TMEM drain, 16 no MMAs, 64 no TMA loads.
for b in 32 64; do
  for mode in 1 5 65 69 9 17 93; do
    echo -n "b=$b mode=$mode : "
    HBSM_F32_MODE=$mode timeout 40 python tools/f32_one.py $b 65536 0.02 1 0 2>&1 | tail -1
    echo
  done
done
