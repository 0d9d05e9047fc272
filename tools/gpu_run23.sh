cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
P="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 6"
for n in 96 128 192; do $P --params '{"variant":"floyd_steinberg"}' --frames $n; done
for n in 64 128; do $P --params '{"variant":"jjn"}' --frames $n; done
$P --params '{"variant":"atkinson"}' --frames 128
