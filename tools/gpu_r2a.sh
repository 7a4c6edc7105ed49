#!/bin/bash
# round 2, first GPU call: (1) full -m gpu suite, (2) K1 A/B of the prepared experiment builds at C2 (+ stage sweep),
# (3) C5 sweep for base / wholebal, (4) PDL A/B on the bench, (5) ncu --set full of lin_bwd and lin_fwd
TAG=${1:-r2a}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log | cut -c1-300
k1() {  # $1 label
  timeout 200 python tools/bench_k1_c2.py --smem 200 --warps 16,32 > gpurun_out/${TAG}_k1_c2_$1.jsonl 2> gpurun_out/${TAG}_k1_c2_$1.err
  echo "$1 k1_c2 rc=$?"; tail -2 gpurun_out/${TAG}_k1_c2_$1.jsonl | cut -c1-330
}
unset DN4GL_LIB; k1 base
for V in whole tworows both balance wholebal; do
  export DN4GL_LIB=$CS/libdn4gl_exp_$V.so; test -f $DN4GL_LIB || continue
  k1 $V
done
for LIBV in base wholebal; do
  if [ $LIBV = base ]; then unset DN4GL_LIB; else export DN4GL_LIB=$CS/libdn4gl_exp_wholebal.so; fi
  for S in 3 4; do
    DN4GL_TILE_STAGES=$S timeout 200 python tools/bench_k1_c2.py --smem 200 --warps 32 > gpurun_out/${TAG}_k1_c2_${LIBV}_stages$S.jsonl 2> gpurun_out/${TAG}_k1_c2_${LIBV}_stages$S.err
    echo "$LIBV stages=$S rc=$?"; tail -1 gpurun_out/${TAG}_k1_c2_${LIBV}_stages$S.jsonl | cut -c1-330
  done
done
for LIBV in base wholebal; do
  if [ $LIBV = base ]; then unset DN4GL_LIB; else export DN4GL_LIB=$CS/libdn4gl_exp_wholebal.so; fi
  timeout 300 python tools/agg_sweep.py --modes tiled --graphs 1024,4096,16384,65536 --dims 32,64,128,256 > gpurun_out/${TAG}_sweep_$LIBV.jsonl 2> gpurun_out/${TAG}_sweep_$LIBV.err
  echo "$LIBV sweep rc=$?"; tail -2 gpurun_out/${TAG}_sweep_$LIBV.jsonl | cut -c1-300
done
unset DN4GL_LIB
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_base.json 2> gpurun_out/${TAG}_bench_base.err
echo "base bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_base.json
for V in 1 2; do
  DN4GL_LIB=$CS/libdn4gl_pdl$V.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_pdl$V.json 2> gpurun_out/${TAG}_bench_pdl$V.err
  echo "pdl$V bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_pdl$V.json
done
timeout 300 python bench.py --no-cpu-baseline --size-hints > gpurun_out/${TAG}_bench_hints.json 2> gpurun_out/${TAG}_bench_hints.err
echo "hints bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_hints.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'lin_bwd_kernel|lin_fwd_kernel' -s 12 -c 4 \
    -o gpurun_out/${TAG}_lin_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_lin.log 2>&1
echo "ncu lin rc=$?"
ls -la gpurun_out | tail -40
