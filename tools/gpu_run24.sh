cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "diffusion or ostro or golden" 2>&1 | tail -4
P="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 6"
for w in 12 16; do for n in 64 128; do echo -n "warps=$w "; DP_WAVE_WARPS=$w $P --params '{"variant":"floyd_steinberg"}' --frames $n; done; done
for w in 12 16; do echo -n "warps=$w "; DP_WAVE_WARPS=$w $P --params '{"variant":"atkinson"}' --frames 128; done
