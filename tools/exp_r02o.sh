#!/bin/bash
# TMA / mbarrier row staging of the lean sampler; estimator with two running products; new host tests
mkdir -p gpurun_out
L=gpurun_out/exp_r02o.log
: > $L
echo "== 16384 / 2048 chains: 131 round-1 | 150 lean | 156 lean+TMA (U8 SG2) | 157 TMA SG4 | 158 TMA U4" >> $L
NKFB_CTA_WARPS=16 TUNE_CHAINS=16384 python tools/tune_sampler.py C4 131,150,156,157,158 2>&1 | grep -v "^MUFU" >> $L
NKFB_CTA_WARPS=16 TUNE_CHAINS=2048 python tools/tune_sampler.py C4 150,156,157 2>&1 | grep -v "^MUFU" >> $L
echo "== step offsets" >> $L
TUNE_STEP_OFFSET=6900 NKFB_CTA_WARPS=16 TUNE_CHAINS=4096 python tools/tune_sampler.py C4 150,156 2>&1 | grep -v "^MUFU" >> $L
cat $L
python -m pytest tests/test_gpu_host.py tests/test_gpu_tensorcore.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5
NKFB_LEAN_TMA=1 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline 2> /dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('bench', d['ms_per_step'], [(k['kernel'][:14], round(k['ms'],3)) for k in d['roofline']['other_kernels']])"
NKFB_LEAN_TMA=1 python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline 2> /dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('bench TMA', d['ms_per_step'], d['roofline']['frac'])"
