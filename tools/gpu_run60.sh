cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench60_$i.json 2> gpurun_out/bench60_$i.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench60_$i.json'))
print($i, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks'])"
done
