#!/bin/bash
# round-2 experiment G: lean sampler with the deferred undo (branch-free decision)
mkdir -p gpurun_out
L=gpurun_out/exp_r02g.log
: > $L
for CH in 16384 4096; do
  echo "== $CH chains: 131 (round-1 CTA kernel), 150-153 (lean, unroll 8/4/2/1)" >> $L
  NKFB_CTA_WARPS=16 TUNE_CHAINS=$CH python tools/tune_sampler.py C4 131,150,151,152,153 2>&1 | grep -v "^MUFU" >> $L
done
NKFB_CTA_WARPS=16 TUNE_CHAINS=4096 ncu --set full --import-source on --clock-control none -k regex:sample_ratio_lean_kernel -c 1 -f \
   -o gpurun_out/prof_r02g_lean python tools/tune_sampler.py C4 153 > gpurun_out/ncu_r02g.log 2>&1
python -m pytest tests/test_gpu_sampler.py -x -q -m gpu > gpurun_out/exp_r02g_tests.log 2>&1
tail -3 gpurun_out/exp_r02g_tests.log
cat $L
