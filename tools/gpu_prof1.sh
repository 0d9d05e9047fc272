set -x
cd $GRAFT_REPO_ROOT
A="python tools/prof_driver.py --mode bayer --params {\"size\":\"8x8\"} --h 1080 --w 1920 --frames 64 --k 16 --reps 3"
B="python tools/prof_driver.py --mode error_diffusion --params {\"variant\":\"floyd_steinberg\"} --h 2160 --w 3840 --frames 8 --k 256 --reps 2"
$A > gpurun_out/profA_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_thresh_tile -s 1 -c 1 -o gpurun_out/prof_thresh_r1a $A > gpurun_out/profA_ncu.log 2>&1
cat gpurun_out/profA_plain.log
$B > gpurun_out/profB_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_diffuse_wave -s 1 -c 1 -o gpurun_out/prof_wave_r1a $B > gpurun_out/profB_ncu.log 2>&1
cat gpurun_out/profB_plain.log
python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/bench_ncu.log
ls -la gpurun_out/
