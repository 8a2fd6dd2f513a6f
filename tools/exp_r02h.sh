#!/bin/bash
# round-2 experiment H: lean sampler (two-pair sub-groups, unrolled row staging); first run of the round-2 bench.py
mkdir -p gpurun_out
L=gpurun_out/exp_r02h.log
: > $L
for CH in 16384; do
  echo "== $CH chains: 131 (round-1 CTA kernel), 150-153 (lean, unroll 8/4/2/1)" >> $L
  NKFB_CTA_WARPS=16 TUNE_CHAINS=$CH python tools/tune_sampler.py C4 131,150,151,152,153 2>&1 | grep -v "^MUFU" >> $L
done
cat $L
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_r02h.err; cat gpurun_out/bench_r02h.json
