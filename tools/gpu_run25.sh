cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "diffusion or ostro or golden or video or strategy" 2>&1 | tail -4
P="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 6"
for n in 1 64 128; do $P --params '{"variant":"floyd_steinberg"}' --frames $n; done
for n in 1 128; do $P --params '{"variant":"jjn"}' --frames $n; done
$P --params '{"variant":"atkinson"}' --frames 128
python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 64 --reps 6 --params '{"variant":"sierra"}' --frames 64
python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --k 64 --reps 6 --frames 64
python tools/prof_driver.py --mode error_diffusion --h 1080 --w 1920 --k 16 --reps 6 --params '{"variant":"floyd_steinberg"}' --frames 256
