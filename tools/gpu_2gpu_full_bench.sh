cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r5c_bench_2gpu.json 2> gpurun_out/r5c_bench_2gpu.err
echo "rc=$?"; grep '^{' gpurun_out/r5c_bench_2gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d.get('strong_scaling'), {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d['configs'].items()})"
grep -i "warn\|error\|Traceback" gpurun_out/r5c_bench_2gpu.err | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r5c_bench_ref_2gpu.json 2> gpurun_out/r5c_bench_ref_2gpu.err
echo "ref rc=$?"; grep '^{' gpurun_out/r5c_bench_ref_2gpu.json | cut -c1-200
