#!/bin/bash
# compute-sanitizer over scripts/sanitize_step.py (round-2 kernels included)
mkdir -p gpurun_out
out=gpurun_out/r2_sanitizer.txt
echo "compute-sanitizer (CUDA 12.9) over scripts/sanitize_step.py on B200, round 2: everything of round 1 plus the second-generation exact row kernel at every CTA size (plain and fused-model instance), the fused training backward, the vector model kernel, the per-output-tile resampling of the multi-CTA path on collapsed weights, the parent-centric gather backward, the multi-warp row statistics" > $out
for tool in memcheck racecheck synccheck; do
  echo "--- $tool" >> $out
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_step.py 2>&1 | grep -E "sanitize driver ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|Traceback|assert| at .*aesmc" | sort | uniq -c | sort -rn | head -12 >> $out
done
cat $out
python scripts/bench_step_variant.py --label default 2>/dev/null | cut -c1-300
