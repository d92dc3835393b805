set -x
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/r02h_multigpu_tests.log 2>&1
echo tests rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-landscape > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
echo bench rc=$?
tail -c 800 gpurun_out/r02h_bench_n2.err
