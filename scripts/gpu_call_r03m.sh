set -x
for k in 24 48 64; do
QAOA_CROSS_CTAS=$k timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 296$k bench.py --gpus 8 --steps 10 --warmup 3 --no-landscape --no-config5 > gpurun_out/r03m_bench_n8_k$k.json 2> gpurun_out/r03m_bench_n8_k$k.err
echo k=$k rc=$?
done
