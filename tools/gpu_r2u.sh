#!/bin/bash
# two-GPU call: NCCL parity test, full -m gpu suite, bench at N = 2 (driver-style launch)
TAG=${1:-r2u}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_zz_nccl_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_nccl.log 2>&1
echo "pytest nccl rc=$?"; tail -6 gpurun_out/${TAG}_pytest_nccl.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest all rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
echo "bench 2gpu rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/${TAG}_bench_2gpu.json; tail -4 gpurun_out/${TAG}_bench_2gpu.err | cut -c1-300
