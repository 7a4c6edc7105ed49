cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench26_2gpu.json 2> gpurun_out/bench26_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/bench26_2gpu.err | cut -c1-300
head -c 400 gpurun_out/bench26_2gpu.json; echo
