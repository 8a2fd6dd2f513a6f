#!/bin/bash
# round-2 experiment D: one chain per warp, up to 32 warps per SM (64 registers) vs two chains per warp, 16 warps
mkdir -p gpurun_out
L=gpurun_out/exp_r02d.log
: > $L
run() { echo "== $1 chains, variant $2, CTA warps $3" >> $L; NKFB_CTA_WARPS=$3 TUNE_CHAINS=$1 python tools/tune_sampler.py C4 $2 2>&1 | grep variant >> $L; }
for CH in 4096 16384; do
  run $CH 131 16
  run $CH 144 32
  run $CH 145 32
  run $CH 147 28
  run $CH 146 24
  run $CH 142 16
done
run 4144 147 28
run 3552 146 24
cat $L
