#!/bin/bash
# First GPU call of a round (≈ 30 min of box time); the K1 experiments are a second call (tools/gpu_ab_k1.sh, ≈ 30 min):
#   make -C dummynode4graphlearning_b200/csrc libdn4gl_pdl1.so libdn4gl_pdl2.so          (here; the .so travel)
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2_first.sh r2a'
# 1. tools/gpu_suite.sh   : pytest -m gpu (incl. the never-run test_zzz_* cases), bench line, ncu launch list, ncu --set
#                           full of the aggregation kernel
# 2. ncu --set full of lin_bwd_kernel (17.5 % of the step's GPU time)
# 3. tools/gpu_ab_pdl.sh  : programmatic dependent launch, parity + bench for both variants
# 3b. bench.py --size-hints (transform without its size read-back)
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bash tools/gpu_suite.sh "$TAG"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lin_bwd_kernel -s 16 -c 2 \
    -o gpurun_out/${TAG}_lin_bwd_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_lin_bwd.log 2>&1
echo "ncu lin_bwd rc=$?"
bash tools/gpu_ab_pdl.sh "$TAG"
# 3b. host-side size hints (no size read-back in the transform), product library
timeout 300 python bench.py --no-cpu-baseline --size-hints > gpurun_out/${TAG}_bench_size_hints.json 2> gpurun_out/${TAG}_bench_size_hints.err
echo "size-hints bench rc=$?"; cut -c1-260 gpurun_out/${TAG}_bench_size_hints.json
ls -la gpurun_out | tail -40
