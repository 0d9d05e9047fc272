set -x
# launch list of the bench command (cold-cache, serialised times: shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-extra --no-video --no-rgb-e2e > gpurun_out/r2k_bench_plain.json 2> gpurun_out/r2k_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-video --no-rgb-e2e > gpurun_out/r2k_ncu_bench.log 2>&1
# the dominant kernel, 128 x 4K, K=256: Floyd-Steinberg and JJN
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_driver.py --mode error_diffusion --params '{"variant":"floyd_steinberg"}' --h 2160 --w 3840 --frames 128 --k 256 --reps 3 > gpurun_out/r2k_fs_plain.log 2>&1 && \
$NCU -k regex:k_diffuse_wave -s 1 -c 1 -f -o gpurun_out/r2k_diffuse_wave_fs_K256_4k_x128 python tools/prof_driver.py --mode error_diffusion --params '{"variant":"floyd_steinberg"}' --h 2160 --w 3840 --frames 128 --k 256 --reps 3 > /dev/null 2>&1
python tools/prof_driver.py --mode error_diffusion --params '{"variant":"jjn"}' --h 2160 --w 3840 --frames 128 --k 256 --reps 3 > gpurun_out/r2k_jjn_plain.log 2>&1 && \
$NCU -k regex:k_diffuse_wave -s 1 -c 1 -f -o gpurun_out/r2k_diffuse_wave_jjn_K256_4k_x128 python tools/prof_driver.py --mode error_diffusion --params '{"variant":"jjn"}' --h 2160 --w 3840 --frames 128 --k 256 --reps 3 > /dev/null 2>&1
python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 16 --reps 3 > gpurun_out/r2k_v4_plain.log 2>&1 && \
$NCU -k regex:k_thresh_v4 -s 1 -c 1 -f -o gpurun_out/r2k_thresh_v4_bayer_K16 python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 16 --reps 3 > /dev/null 2>&1
python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 256 --reps 3 > gpurun_out/r2k_v4w_plain.log 2>&1 && \
$NCU -k regex:k_thresh_v4 -s 1 -c 1 -f -o gpurun_out/r2k_thresh_v4w_bayer_K256 python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 256 --reps 3 > /dev/null 2>&1
cat gpurun_out/r2k_*_plain.log
ls -la gpurun_out/r2k*
