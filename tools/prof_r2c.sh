set -x
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_kmeans.py --frames 16 --k 16 --reps 3 > gpurun_out/r2c_kmeans_plain.log 2>&1 && \
$NCU -k regex:k_kmeans_accum16 -s 1 -c 1 -f -o gpurun_out/r2c_kmeans_accum16 python tools/prof_kmeans.py --frames 16 --k 16 --reps 3 > gpurun_out/r2c_ncu1.log 2>&1
python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 256 --reps 3 > gpurun_out/r2c_fast256_plain.log 2>&1 && \
$NCU -k regex:k_thresh_fast -s 1 -c 1 -f -o gpurun_out/r2c_thresh_fast_bayer_K256 python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 256 --reps 3 > gpurun_out/r2c_ncu2.log 2>&1
python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --frames 38 --k 64 --reps 3 > gpurun_out/r2c_ostro_plain.log 2>&1 && \
$NCU -k regex:k_diffuse_wave -s 1 -c 1 -f -o gpurun_out/r2c_diffuse_wave_ostro_K64_4k_x38 python tools/prof_driver.py --mode ostromoukhov --h 2160 --w 3840 --frames 38 --k 64 --reps 3 > gpurun_out/r2c_ncu3.log 2>&1
python tools/prof_driver.py --mode halftone --h 1080 --w 1920 --frames 64 --k 16 --reps 3 > gpurun_out/r2c_ht_plain.log 2>&1 && \
$NCU -k regex:k_ht_ -s 4 -c 3 -f -o gpurun_out/r2c_halftone_1080p_x64 python tools/prof_driver.py --mode halftone --h 1080 --w 1920 --frames 64 --k 16 --reps 3 > gpurun_out/r2c_ncu4.log 2>&1
python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 16 --reps 3 > gpurun_out/r2c_v4_plain.log 2>&1 && \
$NCU -k regex:k_thresh_v4 -s 1 -c 1 -f -o gpurun_out/r2c_thresh_v4_bayer_K16 python tools/prof_driver.py --mode bayer --params '{"size":"8x8"}' --h 1080 --w 1920 --frames 64 --k 16 --reps 3 > gpurun_out/r2c_ncu5.log 2>&1
cat gpurun_out/r2c_*_plain.log
ls -la gpurun_out/*.ncu-rep
