#!/bin/bash
# round-2 experiment E: lean CTA-synchronous sampler (sampler_lean.cuh) vs the round-1 kernel; unroll factors
mkdir -p gpurun_out
L=gpurun_out/exp_r02e.log
: > $L
for CH in 16384 4096; do
  echo "== $CH chains: variants 131 (round-1 CTA kernel), 150-153 (lean, unroll 8/4/2/1), 154 (lean, one log per 7 pairs)" >> $L
  NKFB_CTA_WARPS=16 TUNE_CHAINS=$CH python tools/tune_sampler.py C4 131,150,151,152,153,154 2>&1 | grep -v "^MUFU" >> $L
done
echo "== step offsets (partial first chunks): 131 vs 151" >> $L
TUNE_STEP_OFFSET=6900 NKFB_CTA_WARPS=16 TUNE_CHAINS=4096 python tools/tune_sampler.py C4 131,151 2>&1 | grep -v "^MUFU" >> $L
cat $L
