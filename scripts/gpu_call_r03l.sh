set -x
python -m pytest tests -m gpu -q > gpurun_out/r03l_gputests.log 2>&1
echo tests rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03l_smoke.log 2>&1
echo smoke rc=$?
