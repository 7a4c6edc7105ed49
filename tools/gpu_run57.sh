cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_pipeline_gpu.py tests/test_mlp_tc_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest57.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest57.log
tail -6 gpurun_out/pytest57.log | cut -c1-300
for i in 1 2; do
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench57_$i.json 2> gpurun_out/bench57_$i.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench57_$i.json'))
print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])"
done
timeout 300 python tools/step_split.py 2>&1 | head -3
