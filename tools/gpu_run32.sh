cd $GRAFT_REPO_ROOT
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1f.json 2>&1; tail -c 300 gpurun_out/bench_ref_r1f.json
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1f.err
python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_ncu.log 2>&1
echo "ncu rc=$?"
