#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp_r02l.log
: > $L
echo "== 150 lean (reductions per chain) | 151 reductions batched SG2 | 152 batched SG4; 16384 and 2048 chains" >> $L
NKFB_CTA_WARPS=16 TUNE_CHAINS=16384 python tools/tune_sampler.py C4 131,150,151,152 2>&1 | grep -v "^MUFU" >> $L
NKFB_CTA_WARPS=16 TUNE_CHAINS=2048 python tools/tune_sampler.py C4 131,150,151,152 2>&1 | grep -v "^MUFU" >> $L
cat $L
