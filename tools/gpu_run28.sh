cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "pixelize or fused or video or golden" 2>&1 | tail -3
P="python tools/prof_driver.py --h 1080 --w 1920 --frames 64 --k 16 --reps 8 --pixelize 270 --upscale 4"
$P --mode blue_noise
$P --mode IGN
$P --mode none
python tools/prof_driver.py --h 1080 --w 1920 --frames 64 --k 16 --reps 8 --pixelize 128 --upscale 1 --mode blue_noise
