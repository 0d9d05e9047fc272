cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
P="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 6"
for v in floyd_steinberg atkinson jjn; do for n in 1 64 128; do $P --params "{\"variant\":\"$v\"}" --frames $n; done; done
