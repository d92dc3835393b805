set -x
python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "chunk" > gpurun_out/r02m_tests.log 2>&1
echo rc=$?
python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/r02m_multigpu_tests.log 2>&1
echo rc=$?
