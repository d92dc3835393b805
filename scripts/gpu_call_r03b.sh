set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r03b_bench_n8.json 2> gpurun_out/r03b_bench_n8.err
echo bench8 rc=$?
python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/r03b_multigpu_tests.log 2>&1
echo tests rc=$?
