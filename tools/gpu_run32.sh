cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest32.log
tail -5 gpurun_out/pytest32.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_finalize_kernel|lin_bwd_reduce_kernel|lin_bwd_kernel|lin_fwd_kernel|bn_act_pool_kernel|bn_bwd_sums_kernel|dot_kernel" -s 60 -c 14 -o gpurun_out/r1d_mlp_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f32.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_f32.log | cut -c1-300
