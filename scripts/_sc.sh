timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_fused_gpu.py -x -q 2>&1 | tail -2
for sc in 1 3 8 20; do for m in exact; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:smc_step --log-file gpurun_out/d.csv python scripts/profile_step.py --mode $m --scale $sc --launches 5 > /dev/null 2>&1
echo "scale=$sc $m: $(grep smc_step gpurun_out/d.csv | tail -1 | awk -F, '{print $NF}') ns"
done; done
