timeout 1200 python -m pytest tests/test_step_parity_gpu.py -x -q 2>&1 | tail -3
AESMC_DEBUG_STATS=1 python scripts/profile_step.py --mode exact --batch 1 --particles 1000000 --launches 1 2>&1 | awk '{ if ($NF+0 > 1000 || /chained/) print }' | tail -20
for b in 1 8 64; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:large_exact_scan --log-file gpurun_out/chain_${b}.csv python scripts/profile_step.py --mode exact --batch $b --particles 1000000 --launches 3 > /dev/null 2>&1
grep large_exact_scan gpurun_out/chain_${b}.csv | tail -1 | awk -F, '{print $(NF-8), $NF}'
done
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | cut -c1-300
