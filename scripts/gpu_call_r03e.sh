set -x
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -k "public_api" > gpurun_out/r03e_multigpu_tests.log 2>&1
echo rc=$?
