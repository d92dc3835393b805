set -x
python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "chunk" > gpurun_out/r02s_tests.log 2>&1
echo rc=$?
python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/r02s_multigpu_tests.log 2>&1
echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-landscape > gpurun_out/r02s_bench_n2.json 2> gpurun_out/r02s_bench_n2.err
echo bench rc=$?
