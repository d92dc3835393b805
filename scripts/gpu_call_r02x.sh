export CUDA_LAUNCH_BLOCKING=1
for cfg in "27 0 1" "27 0 2" "28 0 1" "28 0 2" "29 0 1" "30 1 1" "30 1 2" "28 1 2"; do
timeout 120 python scripts/col_debug.py $cfg
done
