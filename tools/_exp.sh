cd $GRAFT_REPO_ROOT
T="python tools/prof_driver.py --h 1080 --w 1920 --frames 64 --reps 8"
for v in base v4 base v4; do
cp tools/_variants/lib_$v.so dither_pie_b200/libditherpie_b200.so
echo "== $v"
$T --k 16 --mode none | tail -1
$T --k 16 --palette c64 --mode none | tail -1
$T --k 256 --mode none | tail -1
$T --k 64 --mode none | tail -1
done
