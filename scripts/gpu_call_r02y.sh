for cfg in "28 0 3" "28 0 5" "30 1 3" "28 1 5"; do
timeout 120 python scripts/col_debug.py $cfg
done
export CUDA_LAUNCH_BLOCKING=1
for cfg in "28 0 3" "28 0 5"; do
timeout 120 python scripts/col_debug.py $cfg
done
