cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_agg_gpu.py tests/test_mlp_tc_gpu.py tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest54.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest54.log
tail -10 gpurun_out/pytest54.log | cut -c1-300
for i in 1 2; do
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench54_$i.json 2> gpurun_out/bench54_$i.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench54_$i.json'))
print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])
ep=d['breakdown']['entry_points']
print({k:ep[k]['avg_us'] for k in ('dn4gl_bn_act_pool_f32','dn4gl_bn_bwd_sums_f32','dn4gl_dot_f32') if k in ep})"
done
