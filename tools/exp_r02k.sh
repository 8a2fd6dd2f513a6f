#!/bin/bash
# balanced short CTAs in the lean sampler: per-rank steps of the G-GPU shards with and without; reference-text test
mkdir -p gpurun_out
L=gpurun_out/exp_r02k.log
: > $L
echo "== default (short CTAs allowed)" >> $L
SHARD_G=8,4,2,1 python tools/shard_sim.py C4 >> $L 2>&1
echo "== NKFB_NOSHORT=1" >> $L
NKFB_NOSHORT=1 SHARD_G=8,4,2 python tools/shard_sim.py C4 >> $L 2>&1
echo "== forced 24 chains per CTA at G=8 (NKFB_CTA_CHAINS=24: 86 + 86 CTAs, two waves)" >> $L
cat $L
python -m pytest tests/test_reference_text.py tests/test_gpu_sampler.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -8
