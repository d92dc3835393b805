set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "column_kernels" > gpurun_out/r03h_tests.log 2>&1
echo rc=$?
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --column 1 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r03h_bench_col1_$i.json 2> gpurun_out/r03h_bench_col1_$i.err
echo rc=$?
done
timeout 300 python bench.py --steps 10 --warmup 3 --column 0 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r03h_bench_col0.json 2> gpurun_out/r03h_bench_col0.err
