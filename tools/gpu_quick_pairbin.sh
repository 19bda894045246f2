# Quick check of a pair-kernel change: the pairbin / full-size / API parity tests and the N = 1e6 timing.
cd /root/repo
python -m pytest tests/test_gpu_pairbin.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
PB_N=1000000 PB_REPS=4 python tools/pb_run.py 2>&1 | tail -4
