cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lim in 0 140 132 124 116 100; do
DN4GL_SM_LIMIT=$lim timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench55_$lim.json 2> gpurun_out/bench55_$lim.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench55_$lim.json'))
print($lim, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['roofline']['avg_launch_us'],1))"
done
