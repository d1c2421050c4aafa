#!/bin/bash
# compute-sanitizer over scripts/sanitize_step.py (round-2 kernels included).
# racecheck runs twice: on the shipped build, where it reports the level-2 hand-off of smc_step_x.cu (chain_publish /
# chain_wait: one 8-byte word per warp, st.relaxed / ld.relaxed, synchronisation BY a data race as racecheck sees it), and
# on build/variants/libaesmc_sync2.so (scripts/build_variants.sh sync2 "-DAESMC_X_CHAIN_SYNC=2": the same kernel with
# that word written and polled by shared-memory atomics, 1.5x slower), which must be hazard-free.
mkdir -p gpurun_out
out=gpurun_out/r2_sanitizer.txt
echo "compute-sanitizer (CUDA 12.9) over scripts/sanitize_step.py on B200, round 2: everything of round 1 plus the second-generation exact row kernel at every CTA size (plain and fused-model instance), the fused training backward, the vector model kernel, the per-output-tile resampling of the multi-CTA path on collapsed weights, the parent-centric gather backward, the multi-warp row statistics" > $out
filter() { grep -E "sanitize driver ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|Traceback|assert| at .*aesmc" | sed 's/+0x[0-9a-f]*//; s/\[[0-9]* hazards\]//' | sort | uniq -c | sort -rn | head -12; }
for tool in memcheck racecheck synccheck; do
  echo "--- $tool (shipped build)" >> $out
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_step.py 2>&1 | filter >> $out
done
if [ -f build/variants/libaesmc_sync2.so ]; then
  echo "--- racecheck (same sources, -DAESMC_X_CHAIN_SYNC=2: the level-2 hand-off through shared-memory atomics)" >> $out
  AESMC_B200_LIB=$PWD/build/variants/libaesmc_sync2.so timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_step.py 2>&1 | filter >> $out
fi
cat $out
