set -x
timeout 900 python -m pytest tests/test_step_parity_gpu.py -x -q -k "multi_cta" 2>&1 | tail -15
for m in exact fast; do for b in 8 64; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/large_${m}_${b}.csv python scripts/profile_step.py --mode $m --batch $b --particles 1000000 --launches 3 > gpurun_out/large_${m}_${b}.log 2>&1
done; done
