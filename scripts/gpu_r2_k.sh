#!/bin/bash
# parity tests of the step kernels + timing of the default build and every build/variants/*.so
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_fused_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
: > gpurun_out/${TAG}_variants.jsonl
python scripts/bench_step_variant.py --label default >> gpurun_out/${TAG}_variants.jsonl 2>gpurun_out/${TAG}_err.log
for f in build/variants/*.so; do
  [ -f "$f" ] && AESMC_B200_LIB=$PWD/$f python scripts/bench_step_variant.py --label $(basename $f .so) >> gpurun_out/${TAG}_variants.jsonl 2>>gpurun_out/${TAG}_err.log
done
python - <<PY
import json
for l in open("gpurun_out/${TAG}_variants.jsonl"):
    d=json.loads(l); print("%-22s us(s=1) %7.2f  us(s=8) %7.2f  frac %.4f  mism %d/%d lse %d/%d gather %s/%s" % (d["label"], d["us_scale_1"], d["us_scale_8"], d["frac_of_6550_scale_1"], d["mismatch_scale_1"], d["mismatch_scale_8"], d["lse_bits_differ_scale_1"], d["lse_bits_differ_scale_8"], d["gather_ok_scale_1"], d["gather_ok_scale_8"]))
PY
tail -3 gpurun_out/${TAG}_err.log
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_step_x -s 4 -c 1 -f -o gpurun_out/${TAG}_x python scripts/profile_step.py --mode exact > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
fi
