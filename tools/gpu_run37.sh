cd $GRAFT_REPO_ROOT
A="python tools/prof_driver.py --mode error_diffusion --params {\"variant\":\"floyd_steinberg\"} --h 2160 --w 3840 --frames 128 --k 256 --reps 3"
$A > gpurun_out/profW_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_diffuse_wave -s 1 -c 1 -o gpurun_out/prof_wave_r1f $A > gpurun_out/profW_ncu.log 2>&1
echo "ncu wave rc=$?"
B="python tools/prof_driver.py --mode bayer --params {\"size\":\"8x8\"} --h 1080 --w 1920 --frames 64 --k 16 --reps 3"
$B > gpurun_out/profT_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_thresh_v4 -s 1 -c 1 -o gpurun_out/prof_thresh_r1f $B > gpurun_out/profT_ncu.log 2>&1
echo "ncu thresh rc=$?"
