cd $GRAFT_REPO_ROOT
A="python tools/prof_driver.py --mode bayer --params {\"size\":\"8x8\"} --h 1080 --w 1920 --frames 64 --k 16 --reps 3"
$A > gpurun_out/profT_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_thresh_v4 -s 1 -c 1 -o gpurun_out/prof_thresh_r1e $A > gpurun_out/profT_ncu.log 2>&1
echo "ncu rc=$?"
