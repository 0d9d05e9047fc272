cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/wave_timing.py floyd_steinberg 1 | tail -4
for v in floyd_steinberg jjn atkinson; do
for n in 1 8 32 64; do
python tools/prof_driver.py --mode error_diffusion --params "{\"variant\":\"$v\"}" --h 2160 --w 3840 --frames $n --k 256 --reps 3 | tail -1
done
done
python tools/prof_driver.py --mode error_diffusion --params '{"variant":"sierra"}' --h 2160 --w 3840 --frames 32 --k 64 --reps 3 | tail -1
python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --frames 32 --k 64 --reps 3 | tail -1
python tools/prof_driver.py --mode error_diffusion --params '{"variant":"floyd_steinberg"}' --h 1080 --w 1920 --frames 128 --k 16 --reps 3 | tail -1
