#!/bin/bash
# round-2 experiment I: lean sampler variants (sub-group size, pairs per logarithm), site-mode statistics tests, smoke
mkdir -p gpurun_out
L=gpurun_out/exp_r02i.log
: > $L
echo "== 16384 chains: 131 (round-1), 150 U8 SG2 | 151 U8 SG1 | 152 U8 SG4 | 153 U4 SG2 | 154 GL7 SG2 | 155 GL7 SG1" >> $L
NKFB_CTA_WARPS=16 TUNE_CHAINS=16384 python tools/tune_sampler.py C4 131,150,151,152,153,154,155 2>&1 | grep -v "^MUFU" >> $L
cat $L
python -m pytest tests/test_gpu_site_modes.py tests/test_gpu_optimizer.py tests/test_reference_text.py -x -q -m gpu > gpurun_out/exp_r02i_tests.log 2>&1
tail -15 gpurun_out/exp_r02i_tests.log
python __graft_entry__.py smoke > gpurun_out/exp_r02i_smoke.log 2>&1; tail -3 gpurun_out/exp_r02i_smoke.log
