# One GPU-box round trip: parity tests, the per-mode timings quoted in DESIGN.md, the benchmark.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
T="python tools/prof_driver.py --h 1080 --w 1920 --frames 64 --k 16 --reps 8"
$T --mode bayer --params '{"size":"8x8"}'
$T --mode none
$T --mode IGN
$T --mode blue_noise
$T --mode halftone
$T --mode blue_noise --pixelize 270 --upscale 4
D="python tools/prof_driver.py --mode error_diffusion --h 2160 --w 3840 --k 256 --reps 6"
for v in floyd_steinberg atkinson jjn; do for n in 1 128; do $D --params "{\"variant\":\"$v\"}" --frames $n; done; done
python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --k 64 --reps 6 --frames 64
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench.json
