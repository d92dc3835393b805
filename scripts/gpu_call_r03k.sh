set -x
timeout 400 python -m pytest tests/test_gpu_zzz_column_kernels.py -m gpu -q -x > gpurun_out/r03k_tests.log 2>&1
echo rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --column 1 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r03k_bench_col1.json 2> gpurun_out/r03k_bench_col1.err
echo rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --column 0 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r03k_bench_col0.json 2> gpurun_out/r03k_bench_col0.err
echo rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --column 1 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r03k_bench_col1b.json 2> gpurun_out/r03k_bench_col1b.err
