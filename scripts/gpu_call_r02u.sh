set -x
python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "sampling_and_cvar or grover or energy_matches" > gpurun_out/r02u_tests.log 2>&1
echo rc=$?
