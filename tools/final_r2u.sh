# round-2 closing run on one GPU (HEAD): tests, smoke, threshold soak, bench (both arms), launch
# list of the bench command, ncu of the nearest-colour kernel with the skewed cell table
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/final_r2u.sh'
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python tools/soak_threshold.py 2>&1 | tail -2
python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2u_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2u_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])
for k,v in d['modes'].items():
    if 'hbm_frac' in v and ('bayer' in k or 'none' in k or 'IGN' in k or 'blue' in k or 'halftone' in k): print(k, round(v['mpx_s']), round(v['hbm_frac'],3), v.get('ms'))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2u_bench_reference.json 2>/dev/null; cut -c1-120 gpurun_out/r2u_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-video --no-rgb-e2e > gpurun_out/r2u_ncu_bench.log 2>&1; echo ncu rc=$?
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_thresh_v4 -s 1 -c 1 -f -o gpurun_out/r2u_thresh_v4_none_K16 python tools/prof_driver.py --mode none --h 1080 --w 1920 --frames 64 --k 16 --reps 3 > /dev/null 2>&1
ls gpurun_out/r2u*
