# Round-2 final record on ONE GPU: tests, probes, full default bench, remaining ncu captures.
cd /root/repo
TAG=${TAG:-r2g}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
PN=40000 python tools/kmat_probe.py > $O/kmat_probe_$TAG.log 2>&1; cat $O/kmat_probe_$TAG.log
python tools/trsv_probe.py > $O/trsv_probe_$TAG.log 2>&1; cat $O/trsv_probe_$TAG.log
python bench.py > $O/bench_1gpu_$TAG.json 2> $O/bench_1gpu_$TAG.err; tail -3 $O/bench_1gpu_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_1gpu_$TAG.json") if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gp", "cfg", "clocks") if k in d})[:4000])
    print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["timed"])
    print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("gp_fit_predict", {}).get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("bench json unreadable:", e)
PY
cap() {  # name, kernel regex, skip, count, env..., command
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt \
      -o $O/prof_${name}_$TAG -f env "$@" > $O/ncu_${name}_$TAG.log 2>&1 || echo "ncu of $name failed"
}
cap trsv_sweeps_40k trsv_sweep_kernel 18 2 python tools/trsv_probe.py
cap kmat_vk         kmat_sym_kernel   13 1 PN=40000 python tools/kmat_probe.py
ls -la $O/*$TAG*.ncu-rep
