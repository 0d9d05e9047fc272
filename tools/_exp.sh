cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "diffusion or ostromoukhov or hybrid or perceptual or adaptive or config2 or config5 or wavefront or golden" 2>&1 | tail -4
D="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --frames 128 --k 256 --reps 4"
for v in floyd_steinberg atkinson jjn; do
echo "$v bytes"; $D --params "{\"variant\":\"$v\"}" | tail -1
echo "$v nobytes"; DP_WAVE_NO_BYTES=1 $D --params "{\"variant\":\"$v\"}" | tail -1
done
python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --frames 38 --k 64 --reps 4 | tail -1
DP_WAVE_NO_BYTES=1 python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --frames 38 --k 64 --reps 4 | tail -1
