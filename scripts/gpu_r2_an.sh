#!/bin/bash
# racecheck of the exact row kernel's level-2 hand-off under each synchronisation flavour + timing of each
mkdir -p gpurun_out
out=gpurun_out/r2an_racecheck.txt
: > $out
for lib in default build/variants/libaesmc_sync0.so build/variants/libaesmc_sync2.so; do
  echo "--- $lib" >> $out
  if [ "$lib" = default ]; then unset AESMC_B200_LIB; else export AESMC_B200_LIB=$PWD/$lib; fi
  timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_step.py 2>&1 | grep -E "sanitize driver ok|RACECHECK SUMMARY|Error|Traceback|assert" | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -8 >> $out
done
unset AESMC_B200_LIB
cat $out
timeout 300 bash scripts/gpu_r2_k.sh r2an2
