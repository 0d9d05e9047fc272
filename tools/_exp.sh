cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "threshold or wide or golden or index_only or config1 or config3 or deferred or pipeline" 2>&1 | tail -2
T="python tools/prof_driver.py --h 1080 --w 1920 --frames 64 --reps 8"
$T --k 256 --mode bayer --params '{"size":"8x8"}' | tail -1
$T --k 256 --mode IGN | tail -1
$T --k 256 --mode blue_noise | tail -1
$T --k 64 --mode bayer --params '{"size":"8x8"}' | tail -1
$T --k 16 --palette c64 --mode bayer --params '{"size":"8x8"}' | tail -1
DP_THRESH_NO_DEFER=1 $T --k 256 --mode bayer --params '{"size":"8x8"}' | tail -1
