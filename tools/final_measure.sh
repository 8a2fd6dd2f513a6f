#!/bin/bash
# Round-end measurements on one B200 (run under gpurun from the repo root):  bash tools/final_measure.sh TAG
# 1. the bench line (no profiler), 2. the ncu launch list of a short run, 3. ncu --set full of the three kernels of a step,
# (compute-sanitizer is closed on this GPU pool -- gpurun_out/sanitizer_r02*.log -- so step 4 of earlier rounds, a memcheck of
#  C1 through the public API (tools/sanitize_c1.py), is no longer run here)
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
cat gpurun_out/bench_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sample_ratio_lean -c 1 -f -o gpurun_out/prof_${TAG}_sampler \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/ncu_full_${TAG}_sampler.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'infid_fwd_tc_kernel|infid_grad_tc_kernel' -c 2 -f -o gpurun_out/prof_${TAG}_tc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/ncu_full_${TAG}_tc.log 2>&1
ls -la gpurun_out/prof_${TAG}_*.ncu-rep
