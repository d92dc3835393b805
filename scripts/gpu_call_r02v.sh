set -x
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -k "oracle" > gpurun_out/r02v_multigpu_tests.log 2>&1
echo rc=$?
