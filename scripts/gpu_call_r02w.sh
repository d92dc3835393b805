set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "column_kernels" > gpurun_out/r02w_tests.log 2>&1
echo rc=$?
timeout 300 python bench.py --steps 5 --warmup 3 --column 1 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r02w_bench_col1.json 2> gpurun_out/r02w_bench_col1.err
echo rc=$?
timeout 300 python bench.py --steps 5 --warmup 3 --column 0 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r02w_bench_col0.json 2> gpurun_out/r02w_bench_col0.err
echo rc=$?
