timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in exact fast; do for b in 64; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/large_${m}_${b}.csv python scripts/profile_step.py --mode $m --batch $b --particles 1000000 --launches 3 > gpurun_out/large_${m}_${b}.log 2>&1
grep -E "large_search" gpurun_out/large_${m}_${b}.csv | tail -1 | awk -F, '{print $5, $NF}'
done; done
