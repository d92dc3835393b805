set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -x -q -k "grover or xy or per_term or plus_param or config4 or multiangle" > gpurun_out/r02c_tests.log 2>&1
echo tests rc=$?
python - > gpurun_out/r02c_config4.json 2> gpurun_out/r02c_config4.err <<'PY'
import json, bench
print(json.dumps(bench.config4_leg(0), indent=1))
PY
echo c4 rc=$?
