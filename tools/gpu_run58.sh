cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transforms_gpu.py -m gpu -q --tb=short -x -k "match" > gpurun_out/pytest58.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest58.log
tail -6 gpurun_out/pytest58.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench58.json 2> gpurun_out/bench58.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench58.json'))
print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])"
