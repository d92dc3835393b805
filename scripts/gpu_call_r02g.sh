set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_symmetric.py -m gpu -q -x -k "multiangle or cot or xmulti or randomised" > gpurun_out/r02g_tests.log 2>&1
echo rc=$?
python - > gpurun_out/r02g_config4.json 2> gpurun_out/r02g_config4.err <<'PY'
import json, bench
print(json.dumps(bench.config4_leg(0), indent=1))
PY
