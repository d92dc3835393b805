set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err
echo ref rc=$?
# n=32 DRAM traffic of the pass kernels (one evaluation = 11 launches), single metric pass
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tile_prog_kernel -c 11 --csv --log-file gpurun_out/r02b_dram_n32.csv python scripts/prof_tile.py 32 5 1 > gpurun_out/r02b_dram_n32.log 2>&1
echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:"elementwise_pass_kernel|xy_tile_kernel" -c 8 -o gpurun_out/r02b_config4_n28 python scripts/prof_config4.py 28 3 grover,xy > gpurun_out/r02b_config4_ncu.log 2>&1
echo ncu2 rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-optimize --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1
echo ncu3 rc=$?
