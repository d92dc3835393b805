set -x
python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r02j_tests.log 2>&1
echo rc=$?
