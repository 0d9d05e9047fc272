cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
P="python tools/prof_driver.py --h 1080 --w 1920 --frames 64 --k 16 --reps 8"
$P --mode bayer --params '{"size":"8x8"}'
$P --mode none
$P --mode IGN
$P --mode blue_noise
$P --mode bayer --params '{"size":"2x2"}'
$P --mode polka_dot
python tools/prof_driver.py --h 2160 --w 3840 --frames 16 --k 16 --reps 8 --mode bayer --params '{"size":"8x8"}'
DP_THRESH_NO_V4=1 $P --mode bayer --params '{"size":"8x8"}'
