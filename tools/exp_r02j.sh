#!/bin/bash
# full GPU test-suite + bench on the production build
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gputests_r02j.log 2>&1; tail -5 gpurun_out/gputests_r02j.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02j.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['f64']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['other_workloads'].items()}, d['cpu_baseline']['value'])
PY
SHARD_G=8,4,2,1 python tools/shard_sim.py C4
