#!/bin/bash
# two GPUs: NCCL / peer-memory parity test, then the bench at N = 2 with and without the peer all-reduce
TAG=${1:-q2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_nccl_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_nccl.log 2>&1
echo "pytest nccl rc=$?"; tail -12 gpurun_out/${TAG}_pytest_nccl.log | cut -c1-400
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  [ $N -gt $NG ] && break
  for mode in peer nccl; do
    if [ $mode = nccl ]; then export DN4GL_PEER_ALLREDUCE=0; else unset DN4GL_PEER_ALLREDUCE; fi
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29510 + N)) bench.py --gpus $N --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_${N}gpu_${mode}.json 2> gpurun_out/${TAG}_bench_${N}gpu_${mode}.err
    echo "N=$N $mode rc=$?"; grep '^{' gpurun_out/${TAG}_bench_${N}gpu_${mode}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d['breakdown'].get('final_loss'))"
    grep -i "warn\|error" gpurun_out/${TAG}_bench_${N}gpu_${mode}.err | head -3
  done
done
