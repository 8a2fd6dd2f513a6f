#!/bin/bash
# round-2 experiment A: register/LDS microbenchmarks; W=14 vs W=16 CTAs of the CTA-synchronous sampler in ONE launch
mkdir -p gpurun_out
./tools/ubench/regbw_bench > gpurun_out/regbw_bench.log 2>&1
export NKFB_RATIO_FORCECTA=1
for W in 16 14 12; do
  for CH in 4096 2048; do
    echo "== single launch, $CH chains, W=$W" >> gpurun_out/exp_r02a.log
    NKFB_CTA_WARPS=$W TUNE_CHAINS=$CH python tools/tune_sampler.py C4 100 >> gpurun_out/exp_r02a.log 2>&1
  done
done
# ncu of the two shapes (4096 chains: 128 CTAs of 16 warps vs 147 CTAs of 14 warps)
for W in 16 14; do
  NKFB_CTA_WARPS=$W TUNE_CHAINS=4096 ncu --set full --clock-control none -k regex:sample_ratio_cta_kernel -c 1 -f \
     -o gpurun_out/prof_r02a_W$W python tools/tune_sampler.py C4 100 > gpurun_out/ncu_r02a_W$W.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
