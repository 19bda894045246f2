# Envelope factorisation: parity tests (TESTS), timing probe, then the full single-GPU bench.
cd /root/repo
TAG=${TAG:-r2p}
O=gpurun_out
python -m pytest ${TESTS:-tests/test_gpu_envelope.py tests/test_gpu_predict.py tests/test_gpu_api.py tests/test_gpu_dense.py tests/test_gpu_golden_r2.py} -x -q > $O/pytest_env_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_env_$TAG.log; tail -25 $O/pytest_env_$TAG.log
timeout 300 python tools/envelope_probe.py > $O/envelope_probe_$TAG.log 2>&1; cat $O/envelope_probe_$TAG.log | tail -20
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log; tail -2 $O/smoke_$TAG.log
SECONDS=0
python bench.py > $O/bench_1gpu_$TAG.json 2> $O/bench_1gpu_$TAG.err; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench_1gpu_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_1gpu_$TAG.json") if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "e2e", "gp", "cfg") if k in d})[:4000])
    print("variance", json.dumps(d["gp_fit_predict"].get("variance"))[:2500])
    print("wall", json.dumps(d["gp_fit_predict"].get("wall_breakdown_s")), json.dumps(d["gp_fit_predict"].get("kernel_breakdown_s")))
except Exception as e:
    print("bench json unreadable:", e)
PY
