#!/bin/bash
# 2-GPU run: bench (headline, one-collective block, other workloads)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02p_2gpu.json 2> gpurun_out/bench_r02p_2gpu.err
echo "rc=$?"; tail -5 gpurun_out/bench_r02p_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02p_2gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'])
print('one_collective', d.get('one_collective'))
print('f64', d['f64']['ms_per_step'] if 'f64' in d else None, {k:(round(v['ms_per_step'],3), v['unit']) for k,v in d.get('other_workloads',{}).items()})
PY
