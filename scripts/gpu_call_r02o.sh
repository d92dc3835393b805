set -x
for k in 16 20 32 40; do
QAOA_CROSS_CTAS=$k timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$k bench.py --gpus 2 --steps 10 --warmup 3 --no-landscape --no-config5 > gpurun_out/r02o_bench_n2_k$k.json 2> gpurun_out/r02o_bench_n2_k$k.err
echo k=$k rc=$?
done
