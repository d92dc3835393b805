set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/r02e_multigpu_tests.log 2>&1
echo tests rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err
echo bench rc=$?
tail -c 1500 gpurun_out/r02e_bench_n2.err
