set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"large_exact_scan|large_search|large_leaf|large_tree" -s 4 -c 4 -o gpurun_out/large_exact64 python scripts/profile_step.py --mode exact --batch 64 --particles 1000000 --launches 2 > gpurun_out/large_exact64_ncu.log 2>&1
