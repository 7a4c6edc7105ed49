set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transforms_gpu.py tests/test_agg_gpu.py -m gpu -x -q > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest5.log
tail -25 gpurun_out/pytest5.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_us_in_step'], d['roofline']['frac'], d['roofline']['cold_l2_launch_us'])"
