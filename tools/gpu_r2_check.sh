# Round-2 GPU validation: parity tests, smoke, a reduced-size bench (flow check), e2e stage timing.
cd /root/repo
TAG=${TAG:-r2b}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown_$TAG.log 2>&1; cat gpurun_out/e2e_breakdown_$TAG.log
python bench.py ${BENCH_ARGS:---npoints 200000 --gp-train 8000 --gp-predict 100000 --steps 2 --cpu-rows 2000} > gpurun_out/bench_small_$TAG.json 2> gpurun_out/bench_small_$TAG.err; tail -5 gpurun_out/bench_small_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_small_$TAG.json"))
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gp", "cfg", "cpu_baseline") if k in d})[:3000])
except Exception as e:
    print("bench json unreadable:", e)
PY
