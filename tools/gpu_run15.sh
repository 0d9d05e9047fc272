cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
P="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 8"
for v in floyd_steinberg atkinson; do
for n in 1 8 32 64; do
$P --params "{\"variant\":\"$v\"}" --frames $n
done; done
for w in 4 8 12; do echo -n "warps=$w "; DP_WAVE_WARPS=$w $P --params '{"variant":"floyd_steinberg"}' --frames 32; done
for sl in 0 1 4; do echo -n "slack=$sl "; DP_WAVE_SLACK=$sl $P --params '{"variant":"floyd_steinberg"}' --frames 32; done
for n in 1 8 32; do $P --params '{"variant":"jjn"}' --frames $n; done
