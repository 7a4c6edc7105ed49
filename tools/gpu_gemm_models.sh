#!/bin/bash
# native GEMM inside the models: the bench line with it (always / above 1e8 MACs) and with the library GEMMs
TAG=${1:-r4k}
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for mode in tc tc1e8 lib; do
  unset DN4GL_GEMM_TC DN4GL_GEMM_MIN_MACS
  if [ $mode = lib ]; then export DN4GL_GEMM_TC=0; fi
  if [ $mode = tc1e8 ]; then export DN4GL_GEMM_MIN_MACS=1e8; fi
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_${mode}.json 2> gpurun_out/${TAG}_bench_${mode}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${mode}.json"))
print("${mode}", round(d["ms_per_step"],4), {k: (round(v["ms_per_step"],3), v.get("cudaMalloc_calls_in_timed_region")) for k, v in d["configs"].items()})
PY
done
