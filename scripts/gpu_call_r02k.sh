set -x
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/r02k_multigpu_tests.log 2>&1
echo tests rc=$?
for pipe in 1 0; do
QAOA_CHUNK_PIPELINE=$pipe timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$pipe bench.py --gpus 2 --steps 10 --warmup 3 --no-landscape > gpurun_out/r02k_bench_n2_pipe$pipe.json 2> gpurun_out/r02k_bench_n2_pipe$pipe.err
echo bench pipe=$pipe rc=$?
done
tail -c 600 gpurun_out/r02k_bench_n2_pipe1.err
