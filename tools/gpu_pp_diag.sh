# pair-by-pair kernel (block forms off): bit 2 of pairbin_fast_paths on / off, plus the block-form equality test
cd /root/repo
python -m pytest tests/test_gpu_pairbin.py -m gpu -x -q 2>&1 | tail -2
for f in 7 3; do echo "== pair-by-pair FAST=$f"; PB_BLOCK_SUMS=0 PB_FAST=$f PB_N=${N:-1000000} PB_REPS=3 python tools/pb_run.py 2>&1 | tail -3; done
