for rep in 1 2 3; do
for cfg in "28 0 3" "30 1 3" "30 1 5"; do
timeout 120 python scripts/col_debug.py $cfg 2>&1 | tail -2 | cut -c1-160
done
done
timeout 300 python bench.py --steps 5 --warmup 3 --column 1 --no-landscape --no-config4 --no-optimize --no-cpu-baseline > gpurun_out/r03a_bench_col1.json 2> gpurun_out/r03a_bench_col1.err
echo bench rc=$?
