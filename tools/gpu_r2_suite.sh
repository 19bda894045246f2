# Full GPU parity suite at HEAD (+ smoke).
cd /root/repo
TAG=${TAG:-r2r}
O=gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log; tail -16 $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log; tail -2 $O/smoke_$TAG.log
