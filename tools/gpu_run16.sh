cd $GRAFT_REPO_ROOT
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r1c.json; tail -5 gpurun_out/bench_r1c.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1c.json 2>&1; tail -c 600 gpurun_out/bench_ref_r1c.json
A="python tools/prof_driver.py --mode error_diffusion --params {\"variant\":\"floyd_steinberg\"} --h 2160 --w 3840 --frames 32 --k 256 --reps 3"
$A > gpurun_out/profW_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_diffuse_wave -s 1 -c 1 -o gpurun_out/prof_wave_r1d $A > gpurun_out/profW_ncu.log 2>&1
echo "ncu rc=$?"
