cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 16 --reps 3
python tools/prof_driver.py --mode none --h 1080 --w 1920 --frames 64 --k 16 --reps 3
python tools/prof_driver.py --mode IGN --h 1080 --w 1920 --frames 64 --k 16 --reps 3
python tools/prof_driver.py --mode blue_noise --h 1080 --w 1920 --frames 64 --k 16 --reps 3
python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 2160 --w 3840 --frames 16 --k 256 --reps 3
A="python tools/prof_driver.py --mode bayer --params {\"size\":\"8x8\"} --h 1080 --w 1920 --frames 64 --k 16 --reps 3"
$A > gpurun_out/profA_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_thresh -s 1 -c 1 -o gpurun_out/prof_thresh_r1d $A > gpurun_out/profA_ncu.log 2>&1
