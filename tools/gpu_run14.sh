cd $GRAFT_REPO_ROOT
P="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 8"
for v in floyd_steinberg atkinson; do
for n in 16 32 64; do
for w in 4 8 12; do
echo -n "warps=$w "; DP_WAVE_WARPS=$w $P --params "{\"variant\":\"$v\"}" --frames $n
done; done; done
for w in 4 8; do echo -n "warps=$w "; DP_WAVE_WARPS=$w $P --params '{"variant":"jjn"}' --frames 32; done
