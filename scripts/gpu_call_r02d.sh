set -x
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02d_tests.log 2>&1
echo tests rc=$?
python - > gpurun_out/r02d_config4.json 2> gpurun_out/r02d_config4.err <<'PY'
import json, bench
print(json.dumps(bench.config4_leg(0), indent=1))
PY
echo c4 rc=$?
