set -x
python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/r02q_multigpu_tests.log 2>&1
echo rc=$?
