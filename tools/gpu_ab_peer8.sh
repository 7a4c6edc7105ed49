cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for mode in peer nccl peer nccl; do
  if [ $mode = nccl ]; then export DN4GL_PEER_ALLREDUCE=0; else unset DN4GL_PEER_ALLREDUCE; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --no-cpu-baseline --no-extras > gpurun_out/r3q_bench_8gpu_${mode}.json 2> gpurun_out/r3q_bench_8gpu_${mode}.err
  echo "N=8 $mode rc=$?"; grep '^{' gpurun_out/r3q_bench_8gpu_${mode}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d['breakdown'].get('final_loss'))"
  grep -i "warn" gpurun_out/r3q_bench_8gpu_${mode}.err | head -2
done
