set -x
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/r02n_multigpu_tests.log 2>&1
echo tests rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02n_bench_n8.json 2> gpurun_out/r02n_bench_n8.err
echo bench8 rc=$?
QAOA_CHUNK_PIPELINE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-landscape > gpurun_out/r02n_bench_n8_nopipe.json 2> gpurun_out/r02n_bench_n8_nopipe.err
echo bench8 nopipe rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 --no-landscape > gpurun_out/r02n_bench_n4.json 2> gpurun_out/r02n_bench_n4.err
echo bench4 rc=$?
tail -c 600 gpurun_out/r02n_bench_n8.err
