#!/bin/bash
# round-2 experiment C: balanced 28-chain CTAs (straight-line step bodies), per-rank step of G-GPU shards
mkdir -p gpurun_out
L=gpurun_out/exp_r02c.log
: > $L
export NKFB_RATIO_FORCECTA=1 NKFB_CTA_WARPS=16
for CFG in "32 0" "28 0" "24 0" "20 0"; do
  set -- $CFG
  echo "== single launch, 2048 chains, W=16, chains per CTA $1, short mode $2" >> $L
  NKFB_CTA_CHAINS=$1 NKFB_CTA_SHORT=$2 TUNE_CHAINS=2048 python tools/tune_sampler.py C4 100 >> $L 2>&1
done
echo "== single launch 16384 chains, full CTAs" >> $L
TUNE_CHAINS=16384 python tools/tune_sampler.py C4 100 >> $L 2>&1
for G in 8 4 2; do
  echo "== shard_sim G=$G, 28 chains per CTA on both streams" >> $L
  NKFB_CTA_CHAINS=28 SHARD_G=$G python tools/shard_sim.py C4 >> $L 2>&1
done
echo "== shard_sim G=8,4,2,1 default dispatch" >> $L
unset NKFB_RATIO_FORCECTA NKFB_CTA_WARPS
SHARD_G=8,4,2,1 python tools/shard_sim.py C4 >> $L 2>&1
cat $L
