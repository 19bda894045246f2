# Envelope look-ahead: parity tests, probe with the look-ahead variants (1 = U2 after U1, 2 = U2 beside U1, 0 = off).
cd /root/repo
TAG=${TAG:-r2t}
O=gpurun_out
python -m pytest tests/test_gpu_envelope.py -x -q > $O/pytest_envla_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_envla_$TAG.log; tail -4 $O/pytest_envla_$TAG.log
for v in 1 2 0; do
  echo "potrf_env_lookahead = $v"
  ENV_LA=$v timeout 300 python tools/envelope_probe.py > $O/envelope_probe_la${v}_$TAG.log 2>&1; grep -E "potrf dense|loglike" $O/envelope_probe_la${v}_$TAG.log
done
