# round-2 closing run on one GPU: tests, smoke, soaks, bench (both arms), launch list, k-means captures
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for s in threshold diffusion halftone weighted; do timeout 400 python tools/soak_$s.py 2>&1 | tail -2; done
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])
print(d.get('kmeans_sharded'))
for k,v in d['modes'].items(): print(k, round(v['mpx_s']), round(v['hbm_frac'],3), v.get('ms'), v.get('cpu_mpx_s'))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r2_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-video --no-rgb-e2e > gpurun_out/r2m_ncu_bench.log 2>&1; echo ncu rc=$?; wc -l gpurun_out/r2m_bench_launches.csv
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 python tools/prof_kmeans.py --frames 16 --k 16 --reps 3 && \
timeout 400 $NCU -k regex:k_kmeans_accum16 -s 1 -c 1 -f -o gpurun_out/r2m_kmeans_accum16 python tools/prof_kmeans.py --frames 16 --k 16 --reps 3 > /dev/null 2>&1
timeout 200 python tools/prof_kmeans.py --frames 1 --k 16 --reps 3 --loop 20 && \
timeout 400 $NCU -k regex:k_kmeans_loop -s 1 -c 1 -f -o gpurun_out/r2m_kmeans_loop_4k_x20 python tools/prof_kmeans.py --frames 1 --k 16 --reps 2 --loop 20 > /dev/null 2>&1
ls -la gpurun_out/r2m* gpurun_out/r2_*
