cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench45_8gpu.json 2> gpurun_out/bench45_8gpu.err; echo "bench8 rc=$?"; tail -5 gpurun_out/bench45_8gpu.err | cut -c1-300
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench45_8gpu.json') if l.startswith('{')][-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['n_gpus'], d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench45_4gpu.json 2> gpurun_out/bench45_4gpu.err; echo "bench4 rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench45_4gpu.json') if l.startswith('{')][-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['n_gpus'])"
