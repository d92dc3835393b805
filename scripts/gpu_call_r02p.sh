set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_symmetric.py -m gpu -q -x > gpurun_out/r02p_tests.log 2>&1
echo rc=$?
python scripts/landscape_batch.py 32 64 128 256 0 > gpurun_out/r02p_landscape_batch.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:"grover_pass_c64_kernel|xy_compact_gate_kernel" -c 10 -o gpurun_out/r02p_config4_new_n28 python scripts/prof_config4.py 28 3 grover,xy > gpurun_out/r02p_ncu.log 2>&1
echo ncu rc=$?
