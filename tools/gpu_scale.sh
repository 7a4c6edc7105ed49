#!/bin/bash
# driver-style scaling run on one box: bench.py at N = 1, 2, 4, 8 (whatever the box has) -> gpurun_out/<tag>_bench_<N>gpu.json
TAG=${1:-scale}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N = 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  fi
  echo "N=$N rc=$?"; grep '^{' gpurun_out/${TAG}_bench_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d.get('strong_scaling'))"
done
