# Full GPU parity suite at HEAD + ncu launch list of one envelope likelihood evaluation (N = 40k).
cd /root/repo
TAG=${TAG:-r2r}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
timeout 200 python tools/envelope_once.py || echo "envelope_once failed"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_envelope_$TAG.csv \
    python tools/envelope_once.py > $O/ncu_launches_envelope_$TAG.log 2>&1 || echo "launch list failed"
wc -l $O/launches_envelope_$TAG.csv
