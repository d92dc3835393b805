for rep in 1 2 3 4; do
for cfg in "28 0 3" "30 1 3"; do
timeout 120 python scripts/col_debug.py $cfg 2>&1 | tail -2 | cut -c1-160
done
done
