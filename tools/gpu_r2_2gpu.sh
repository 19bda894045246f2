# 2-GPU record of the final build (driver's launch line).
cd /root/repo
TAG=${TAG:-r2s}
O=gpurun_out
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NG:-2} \
    > $O/bench_${NG:-2}gpu_$TAG.json 2> $O/bench_${NG:-2}gpu_$TAG.err; echo "bench rc=$? wall=${SECONDS}s"; tail -5 $O/bench_${NG:-2}gpu_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${NG:-2}gpu_$TAG.json") if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gp", "cfg", "clocks") if k in d})[:4000])
except Exception as e:
    print("bench json unreadable:", e)
PY
