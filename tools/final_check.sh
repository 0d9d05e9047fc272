cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r1n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1n.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])
for k,v in d['modes'].items(): print(k, round(v['mpx_s']), round(v['hbm_frac'],3), v.get('write_frac'), v.get('cpu_mpx_s'))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1n.json 2>/dev/null; cut -c1-400 gpurun_out/bench_ref_r1n.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/ncu_r1n.log 2>&1; echo ncu rc=$?; wc -l gpurun_out/launches_r1n.csv
