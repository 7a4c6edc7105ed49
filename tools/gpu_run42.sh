cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do
DN4GL_BENCH_TRACE=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench42_$i.json 2> gpurun_out/bench42_$i.err
python -c "
import json
d=json.load(open('gpurun_out/bench42_$i.json'))
print($i, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['samples'])"
grep "e2e submit" gpurun_out/bench42_$i.err | tr ' ' '\n' | awk '$1+0>2.5' | tr '\n' ' '; echo " <- submit spikes"
grep "e2e result" gpurun_out/bench42_$i.err | tr ' ' '\n' | awk '$1+0>2.5' | tr '\n' ' '; echo " <- result spikes"
done
