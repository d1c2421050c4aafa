#!/bin/bash
# time the default build and every build/variants/*.so (bench_step_variant.py), optional parity tests first
# usage: gpu_variants.sh TAG [pytest]
TAG=$1
mkdir -p gpurun_out
if [ "$2" = "pytest" ]; then
  timeout 900 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
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
