set -x
python scripts/landscape_batch.py > gpurun_out/r02f_landscape_batch.jsonl 2>&1
python - > gpurun_out/r02f_config4.json 2> gpurun_out/r02f_config4.err <<'PY'
import json, bench
print(json.dumps(bench.config4_leg(0), indent=1))
PY
python -m pytest tests/test_gpu_znotebooks.py -m gpu -q -x -k cvar > gpurun_out/r02f_tests.log 2>&1
echo rc=$?
