#!/bin/bash
# One gpurun call that produces what profiles/ needs for a round (run from the repo root on the GPU box):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r3a'                 one GPU
#   gpurun --gpus 2 --timeout 2400 -- 'bash tools/gpu_r2u.sh r3a'          two GPUs: NCCL parity test + bench at N = 2
# 1. pytest -m gpu (writes gpurun_out/parity_errors.json)   2. full bench line + reference arm   3. smoke()
# 4. ncu --set full of the hot kernels (3 launches each) -> tools/ncu_summary.py
# Optional: per-role wait accounting of the pipelined stage kernels --
#   make -C dummynode4graphlearning_b200/csrc libdn4gl_exp.so EXP_FLAGS=-DDN4GL_PIPE_TL && mv .../libdn4gl_exp.so .../libdn4gl_pipetl.so
#   DN4GL_LIB=$PWD/dummynode4graphlearning_b200/csrc/libdn4gl_pipetl.so python tools/pipe_timeline.py
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "bench ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lin_fwd_pipe|lin_bwd_pipe|spmm_pipe' -s 24 -c 9 \
    -o gpurun_out/${TAG}_hot python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_hot.log 2>&1
echo "ncu hot rc=$?"
