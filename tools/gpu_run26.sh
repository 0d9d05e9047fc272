cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "diffusion or ostro" 2>&1 | tail -3
P="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 6"
for l in 0 1; do for v in floyd_steinberg atkinson jjn; do echo -n "l1smem=$l "; DP_WAVE_L1SMEM=$l $P --params "{\"variant\":\"$v\"}" --frames 128; done; done
for l in 0 1; do echo -n "l1smem=$l "; DP_WAVE_L1SMEM=$l $P --params '{"variant":"floyd_steinberg"}' --frames 64; done
echo default; $P --params '{"variant":"floyd_steinberg"}' --frames 128; $P --params '{"variant":"atkinson"}' --frames 128
