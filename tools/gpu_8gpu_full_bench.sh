#!/bin/bash
# final check on one 8-GPU box: the driver-style launch of the bench with every config at N = 8
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r5d_bench_8gpu.json 2> gpurun_out/r5d_bench_8gpu.err
echo "rc=$?"; grep '^{' gpurun_out/r5d_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d.get('strong_scaling'), {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d['configs'].items()})"
grep -i "warn\|error\|Traceback" gpurun_out/r5d_bench_8gpu.err | head -5
