cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for rep in 1 2; do
for mode in coop nocoop; do
  if [ $mode = nocoop ]; then export DN4GL_NO_COOP=1; else unset DN4GL_NO_COOP; fi
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/ab_${mode}_${rep}.json 2> gpurun_out/ab_${mode}_${rep}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_${mode}_${rep}.json"))
print("${mode} ${rep}", round(d["ms_per_step"],4), round(d["l2_flush"]["ms_per_step_without_flush"],4), {k: (round(v["ms_per_step"],3), v.get("cudaMalloc_calls_in_timed_region")) for k, v in d["configs"].items()})
PY
done; done
