# Round-2 closing record on ONE GPU at HEAD (shared-geometry bootstrap is the default): parity tests, smoke, full
# default bench.  Outputs under gpurun_out/ with tag $TAG.  TESTS selects the pytest arguments.
cd /root/repo
TAG=${TAG:-r2n}
O=gpurun_out
python -m pytest ${TESTS:-tests -m gpu} -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log; tail -15 $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log; tail -2 $O/smoke_$TAG.log
SECONDS=0
python bench.py > $O/bench_1gpu_$TAG.json 2> $O/bench_1gpu_$TAG.err; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench_1gpu_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_1gpu_$TAG.json") if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gp", "cfg", "clocks") if k in d})[:5000])
    print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["timed"]["ms_per_launch"])
    print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("gp_fit_predict", {}).get("cpu_baseline", {}).get("value"))
    print("variance", json.dumps(d["gp_fit_predict"].get("variance"))[:2500])
except Exception as e:
    print("bench json unreadable:", e)
PY
