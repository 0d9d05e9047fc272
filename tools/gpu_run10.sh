cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in floyd_steinberg jjn atkinson; do
python tools/prof_driver.py --mode error_diffusion --params "{\"variant\":\"$v\"}" --h 2160 --w 3840 --frames 1 --k 256 --reps 3
python tools/prof_driver.py --mode error_diffusion --params "{\"variant\":\"$v\"}" --h 2160 --w 3840 --frames 8 --k 256 --reps 3
done
python tools/prof_driver.py --mode error_diffusion --params '{"variant":"floyd_steinberg"}' --h 2160 --w 3840 --frames 32 --k 256 --reps 3
python tools/prof_driver.py --mode error_diffusion --params '{"variant":"sierra"}' --h 2160 --w 3840 --frames 8 --k 64 --reps 3
python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --frames 8 --k 64 --reps 3
