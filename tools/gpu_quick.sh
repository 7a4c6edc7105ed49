#!/bin/bash
# short GPU call: selected tests, the C2 bench line, (optionally) the per-role timeline of the stage kernels
TAG=${1:-q}
SEL=${2:-tests/test_agg_gpu.py tests/test_models_gpu.py tests/test_mlp_tc_gpu.py}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest $SEL -m gpu -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d.get("l2_flush"))
PY
tail -3 gpurun_out/${TAG}_bench.err
if [ -f dummynode4graphlearning_b200/csrc/libdn4gl_pipetl.so ]; then
  DN4GL_LIB=$PWD/dummynode4graphlearning_b200/csrc/libdn4gl_pipetl.so timeout 300 python tools/pipe_timeline.py --rows 156759 --dims 32 > gpurun_out/${TAG}_timeline.txt 2>&1
  grep '"op": "bwd' gpurun_out/${TAG}_timeline.txt | cut -c1-1500
fi
