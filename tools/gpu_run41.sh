cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
DN4GL_BENCH_TRACE=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench41_$i.json 2> gpurun_out/bench41_$i.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench41_$i.json'))
print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['samples'], d['breakdown']['cudaMalloc_calls_in_timed_region'], d['breakdown']['cudaMalloc_calls_in_e2e_region'])"
grep "host ms" gpurun_out/bench41_$i.err | cut -c1-400
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv; nproc; cat /proc/cpuinfo | grep "model name" | head -1
