set -x
cd /root/repo
python -m pytest tests/test_gpu_pairbin.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/pytest_gpu_r1e.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1e.log
for f in 0 1 2 3; do echo "FAST=$f"; PB_FAST=$f PB_N=1000000 PB_REPS=3 python tools/pb_run.py; done > gpurun_out/pb_fast_r1j.log 2>&1
for f in 0 3; do echo "W FAST=$f"; PB_W=1 PB_FAST=$f PB_N=400000 PB_REPS=3 python tools/pb_run.py; done >> gpurun_out/pb_fast_r1j.log 2>&1
cat gpurun_out/pb_fast_r1j.log
python bench.py --skip-gp --skip-cpu --skip-other --steps 5 --warmup 3 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1j.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['timed_kernel'])"
