set -x
timeout 500 python -m pytest tests/test_gpu_znotebooks.py tests/test_gpu_api.py tests/test_gpu_zflip.py -m gpu -q -x -k "cvar or CVaR or exactcover or flip" > gpurun_out/r03n_tests.log 2>&1
echo rc=$?
