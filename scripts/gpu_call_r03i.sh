set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "column_kernels" > gpurun_out/r03i_tests.log 2>&1
echo rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --column 1 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r03i_bench_col1.json 2> gpurun_out/r03i_bench_col1.err
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"column_prog_kernel" -c 3 -o gpurun_out/r03i_column_n28 python scripts/prof_column.py 28 3 > gpurun_out/r03i_ncu.log 2>&1
echo ncu rc=$?
