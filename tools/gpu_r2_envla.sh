# Envelope look-ahead: parity tests, probe with the look-ahead on and off.
cd /root/repo
TAG=${TAG:-r2t}
O=gpurun_out
python -m pytest tests/test_gpu_envelope.py tests/test_gpu_predict.py -x -q > $O/pytest_envla_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_envla_$TAG.log; tail -8 $O/pytest_envla_$TAG.log
timeout 300 python tools/envelope_probe.py > $O/envelope_probe_$TAG.log 2>&1; cat $O/envelope_probe_$TAG.log | tail -14
ENV_LA=0 timeout 300 python tools/envelope_probe.py > $O/envelope_probe_nola_$TAG.log 2>&1; grep -E "potrf dense|loglike" $O/envelope_probe_nola_$TAG.log
