#!/bin/bash
# A/B of the experimental aggregation-kernel builds against the product build, one gpurun call:
#   make -C dummynode4graphlearning_b200/csrc libdn4gl_exp_whole.so libdn4gl_exp_tworows.so libdn4gl_exp_balance.so libdn4gl_exp_wholebal.so
#   (here, before the call: the .so travel; `make libdn4gl_exp.so EXP_FLAGS=...` for anything else)
#   gpurun --timeout 2400 -- 'bash tools/gpu_ab_k1.sh r2a'      (≈ 5 min per library: parity 2, C2 1, sweep 2-3)
# For every libdn4gl_exp*.so present: 1. parity (the aggregation tests through DN4GL_LIB), 2. K1 alone on the C2
# structure and the C5 sweep; the same two measurements for the product build first.  Then 3 / 4 smaller stages with the
# product build.  Outputs under gpurun_out/<tag>_ab_*.
#   whole   = -DDN4GL_TILE_WHOLE_GRAPHS  graphs that span a window but fit a stage become their own closed tile
#             (tools/k1_tiles_model.py: rows on the checked slow path at C2 15.2 % -> 5.8 %)
#   tworows = -DDN4GL_K1_TWO_ROWS        two rows per sub-group in flight (process_tile_fast2)
#   balance = -DDN4GL_TILE_BALANCE       tile descriptors permuted so that the round-robin deal gives every CTA about the
#             same estimated work (model: max / mean CTA load 1.7-2.2 -> 1.12-1.25); wholebal = whole + balance
#   (measured in round 1, no gain: -DDN4GL_K1_TAIL_ILP)
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
run_pair() {   # $1 = label; DN4GL_LIB set by the caller (or unset for the product build)
  timeout 300 python tools/bench_k1_c2.py --smem 200 --warps 16,32 > gpurun_out/${TAG}_ab_k1_c2_$1.jsonl 2> gpurun_out/${TAG}_ab_k1_c2_$1.err
  echo "$1 k1_c2 rc=$?"; tail -3 gpurun_out/${TAG}_ab_k1_c2_$1.jsonl | cut -c1-300
  timeout 400 python tools/agg_sweep.py > gpurun_out/${TAG}_ab_sweep_$1.jsonl 2> gpurun_out/${TAG}_ab_sweep_$1.err
  echo "$1 sweep rc=$?"; tail -3 gpurun_out/${TAG}_ab_sweep_$1.jsonl | cut -c1-300
}
unset DN4GL_LIB
run_pair base
for EXP in $CS/libdn4gl_exp*.so; do
  test -f "$EXP" || continue
  V=$(basename $EXP .so); V=${V#libdn4gl_}
  export DN4GL_LIB=$EXP
  timeout 600 python -m pytest tests/test_agg_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_ab_pytest_$V.log 2>&1
  echo "$V parity rc=$?"; tail -2 gpurun_out/${TAG}_ab_pytest_$V.log | cut -c1-200
  run_pair $V
done
# three / four smaller stages instead of the automatic choice (2 at C2): product build and, if present, whole + balance
# (the host model predicts that smaller stages need the tile experiments: DESIGN.md section 9 item 1)
for LIBV in base wholebal; do
  if [ $LIBV = base ]; then unset DN4GL_LIB; else test -f $CS/libdn4gl_exp_wholebal.so || continue; export DN4GL_LIB=$CS/libdn4gl_exp_wholebal.so; fi
  for S in 3 4; do
    DN4GL_TILE_STAGES=$S timeout 300 python tools/bench_k1_c2.py --smem 200 --warps 16,32 > gpurun_out/${TAG}_ab_k1_c2_${LIBV}_stages$S.jsonl 2> gpurun_out/${TAG}_ab_k1_c2_${LIBV}_stages$S.err
    echo "$LIBV stages=$S rc=$?"; tail -3 gpurun_out/${TAG}_ab_k1_c2_${LIBV}_stages$S.jsonl | cut -c1-300
  done
done
unset DN4GL_LIB
