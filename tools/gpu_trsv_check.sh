cd /root/repo
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_predict.py tests/test_gpu_api.py -x -q 2>&1 | tail -8
timeout 300 python tools/trsv_probe.py 2>&1 | tail -5
