set -x
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02r_gputests.log 2>&1
echo tests rc=$?
python bench.py --steps 10 --warmup 3 > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err
echo bench rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1
echo smoke rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02r_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-optimize --no-cpu-baseline > gpurun_out/r02r_ncu_bench.log 2>&1
echo ncu rc=$?
