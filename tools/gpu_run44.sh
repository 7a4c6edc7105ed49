cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench44_2gpu.json 2> gpurun_out/bench44_2gpu.err; echo "bench2 rc=$?"; tail -5 gpurun_out/bench44_2gpu.err | cut -c1-300
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench44_2gpu.json') if l.startswith('{')][-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['n_gpus'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench44_ref2.json 2> gpurun_out/bench44_ref2.err; echo "ref2 rc=$?"; cut -c1-200 gpurun_out/bench44_ref2.json
