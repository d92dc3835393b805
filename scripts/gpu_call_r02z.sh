for sync in 0 1 2 3; do
for cfg in "28 0 3" "30 1 3"; do
echo "sync=$sync"
timeout 120 python scripts/col_debug.py $cfg $sync 2>&1 | tail -3 | cut -c1-200
done
done
