cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py 2>&1 | grep -v "^W\|warn" | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra --batch 32 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; tail -c 900 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python bench.py --gpus 1 --steps 3 --warmup 3 --no-extra --batch 32 > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench1 rc=$?"; tail -c 900 gpurun_out/bench_n1b.json
