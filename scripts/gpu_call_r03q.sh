set -x
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r03q_gputests.log 2>&1
echo tests rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03q_smoke.log 2>&1
echo smoke rc=$?
python bench.py --steps 10 --warmup 3 > gpurun_out/r03q_bench_n1.json 2> gpurun_out/r03q_bench_n1.err
echo bench rc=$?
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r03q_bench_ref.json 2> gpurun_out/r03q_bench_ref.err
echo ref rc=$?
