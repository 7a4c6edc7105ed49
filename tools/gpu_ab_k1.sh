#!/bin/bash
# A/B of the experimental aggregation-kernel build against the product build, one gpurun call:
#   make -C dummynode4graphlearning_b200/csrc libdn4gl_exp.so [EXP_FLAGS=-DDN4GL_K1_TWO_ROWS]   (here, before the call: the
#   .so travels; default EXP_FLAGS = -DDN4GL_K1_TAIL_ILP, measured in round 1: no gain)
#   gpurun --timeout 900 -- 'bash tools/gpu_ab_k1.sh r2a'
# 1. parity of the experimental build (the aggregation tests through DN4GL_LIB), 2. K1 alone on the C2 structure and
# the C5 sweep with both builds.  Outputs under gpurun_out/<tag>_ab_*.
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
EXP=$PWD/dummynode4graphlearning_b200/csrc/libdn4gl_exp.so
test -f "$EXP" || { echo "build libdn4gl_exp.so first"; exit 1; }
DN4GL_LIB=$EXP timeout 600 python -m pytest tests/test_agg_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_ab_pytest_exp.log 2>&1
echo "exp parity rc=$?"; tail -2 gpurun_out/${TAG}_ab_pytest_exp.log | cut -c1-200
for V in base exp; do
  if [ $V = exp ]; then export DN4GL_LIB=$EXP; else unset DN4GL_LIB; fi
  timeout 300 python tools/bench_k1_c2.py > gpurun_out/${TAG}_ab_k1_c2_$V.jsonl 2> gpurun_out/${TAG}_ab_k1_c2_$V.err
  echo "$V k1_c2 rc=$?"; tail -3 gpurun_out/${TAG}_ab_k1_c2_$V.jsonl | cut -c1-300
  timeout 400 python tools/agg_sweep.py > gpurun_out/${TAG}_ab_sweep_$V.jsonl 2> gpurun_out/${TAG}_ab_sweep_$V.err
  echo "$V sweep rc=$?"; tail -3 gpurun_out/${TAG}_ab_sweep_$V.jsonl | cut -c1-300
done
unset DN4GL_LIB
for S in 3 4; do   # three / four smaller stages instead of the automatic choice (2 at C2), product build
  DN4GL_TILE_STAGES=$S timeout 300 python tools/bench_k1_c2.py > gpurun_out/${TAG}_ab_k1_c2_stages$S.jsonl 2> gpurun_out/${TAG}_ab_k1_c2_stages$S.err
  echo "stages=$S rc=$?"; tail -3 gpurun_out/${TAG}_ab_k1_c2_stages$S.jsonl | cut -c1-300
done
