# Full GPU validation of the tree (run through gpurun): parity tests, smoke, default bench; TAG names the outputs.
cd /root/repo
TAG=${TAG:-r1m}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json")); g = d["gp_fit_predict"]; r = d["roofline"]
print("2pcf", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "pp frac", r["frac"], r["ms_per_launch"], d["clocks"])
print("timed", r["timed_kernel"])
print("gp", g["value"], g["wall_breakdown_s"], g["roofline_potrf"]["frac"], g["roofline_kmat"]["frac"])
o = d["other_configs"]
print("other", o["loglike_fit_N10k"]["fit_wall_s"], o["bootstrap_2pcf_N200k"]["bootstrap_wall_s"], "cpu", d["cpu_baseline"]["value"])
PY
