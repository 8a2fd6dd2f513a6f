#!/bin/bash
# round-2 experiment F: ncu of the lean sampler (unroll 1), sampler tests with the lean kernel as the default
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sampler.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/exp_r02f_tests.log 2>&1
tail -5 gpurun_out/exp_r02f_tests.log
NKFB_CTA_WARPS=16 TUNE_CHAINS=4096 ncu --set full --import-source on --clock-control none -k regex:sample_ratio_lean_kernel -c 1 -f \
   -o gpurun_out/prof_r02f_lean python tools/tune_sampler.py C4 153 > gpurun_out/ncu_r02f.log 2>&1
SHARD_G=8,4,2,1 python tools/shard_sim.py C4 > gpurun_out/exp_r02f_shard.log 2>&1
cat gpurun_out/exp_r02f_shard.log
