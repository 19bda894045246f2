# Round-2 record on ONE GPU: full default bench, dense probe, ncu captures of the final kernels, launch list.
cd /root/repo
TAG=${TAG:-r2f}
O=gpurun_out
python bench.py > $O/bench_1gpu_$TAG.json 2> $O/bench_1gpu_$TAG.err; tail -3 $O/bench_1gpu_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("$O/bench_1gpu_$TAG.json"))
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gp", "cfg", "clocks") if k in d})[:4000])
    print("roofline", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["roofline"]["timed"]["ms_per_launch"])
    print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("gp_fit_predict", {}).get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("bench json unreadable:", e)
PY
DP_N=4096,10000,20000 python tools/dense_probe.py > $O/dense_probe_$TAG.log 2>&1; grep potrf $O/dense_probe_$TAG.log
cap() {  # name, kernel regex, skip, count, env..., command
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt \
      -o $O/prof_${name}_$TAG -f env "$@" > $O/ncu_${name}_$TAG.log 2>&1 || echo "ncu of $name failed"
}
cap pairbin_timed   pairbin_kernel     1 1 PB_N=1000000 PB_REPS=2 python tools/pb_run.py
cap pairbin_witness pairbin_kernel     1 1 PB_BLOCK_SUMS=0 PB_N=1000000 PB_REPS=2 python tools/pb_run.py
cap trsv_sweeps     trsv_sweep_kernel  4 2 python tools/trsv_probe.py
cap trsv_prep       trsv_prep_kernel   2 1 python tools/trsv_probe.py
python bench.py --skip-gp --skip-cpu --skip-other --npoints 200000 --steps 2 --warmup 3 > $O/bench_200k_$TAG.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --skip-gp --skip-cpu --skip-other --npoints 200000 --steps 2 --warmup 3 > $O/ncu_launch_$TAG.log 2>&1
wc -l $O/launches_$TAG.csv; ls -la $O/*$TAG*.ncu-rep
