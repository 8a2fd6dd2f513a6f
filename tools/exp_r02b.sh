#!/bin/bash
# round-2 experiment B: 16-warp CTAs carrying 28 chains (one chain short on 4 warps) vs full CTAs; which warps share a sub-partition
mkdir -p gpurun_out
export NKFB_RATIO_FORCECTA=1
L=gpurun_out/exp_r02b.log
: > $L
for CFG in "32 0" "28 0" "28 1" "24 0" "30 0"; do
  set -- $CFG
  echo "== single launch, 4096 chains, W=16, chains per CTA $1, short mode $2" >> $L
  NKFB_CTA_WARPS=16 NKFB_CTA_CHAINS=$1 NKFB_CTA_SHORT=$2 TUNE_CHAINS=4096 python tools/tune_sampler.py C4 100 >> $L 2>&1
done
echo "== 4144 chains = 148 x 28, W=16, 28 per CTA" >> $L
NKFB_CTA_WARPS=16 NKFB_CTA_CHAINS=28 NKFB_CTA_SHORT=0 TUNE_CHAINS=4144 python tools/tune_sampler.py C4 100 >> $L 2>&1
echo "== 16384 chains, default vs 28 per CTA" >> $L
NKFB_CTA_WARPS=16 TUNE_CHAINS=16384 python tools/tune_sampler.py C4 100 >> $L 2>&1
python -m pytest tests/test_gpu_sampler.py -x -q -m gpu > gpurun_out/exp_r02b_tests.log 2>&1
tail -3 gpurun_out/exp_r02b_tests.log
cat $L
