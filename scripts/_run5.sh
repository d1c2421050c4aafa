timeout 1200 python -m pytest tests/test_step_parity_gpu.py -x -q -k multi_cta 2>&1 | tail -3
for t in 256 512 1024; do for b in 8 64; do
AESMC_SPAN_THREADS=$t timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:large_exact_scan --log-file gpurun_out/chain_${b}.csv python scripts/profile_step.py --mode exact --batch $b --particles 1000000 --launches 3 > /dev/null 2>&1
echo "NT=$t B=$b K=1e6 $(grep large_exact_scan gpurun_out/chain_${b}.csv | tail -1 | awk -F, '{print $NF}')"
AESMC_SPAN_THREADS=$t timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:large_exact_scan --log-file gpurun_out/chain_${b}.csv python scripts/profile_step.py --mode exact --batch $b --particles 100000 --launches 3 > /dev/null 2>&1
echo "NT=$t B=$b K=1e5 $(grep large_exact_scan gpurun_out/chain_${b}.csv | tail -1 | awk -F, '{print $NF}')"
done; done
