cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_driver.py --h 1080 --w 1920 --frames 64 --k 16 --reps 8 --mode halftone
python tools/prof_driver.py --h 2160 --w 3840 --frames 16 --k 16 --reps 8 --mode halftone
python tools/prof_driver.py --h 1080 --w 1920 --frames 1 --k 16 --reps 8 --mode halftone
