AESMC_DEBUG_STATS=1 python scripts/profile_step.py --mode exact --batch 8 --particles 1000000 --launches 2 2>&1 | tail -4
AESMC_DEBUG_STATS=1 python scripts/profile_step.py --mode exact --batch 1 --particles 1000000 --launches 2 2>&1 | tail -4
for b in 1 2 8; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:large_exact_scan --log-file gpurun_out/chain_${b}.csv python scripts/profile_step.py --mode exact --batch $b --particles 1000000 --launches 3 > /dev/null 2>&1
grep large_exact_scan gpurun_out/chain_${b}.csv | tail -1
done
