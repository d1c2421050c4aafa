#!/bin/bash
# guarded first run of a restructured step kernel: parity tests under a short timeout, then the variant timing
TAG=${1:-r2am}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_step_parity_gpu.py tests/test_fused_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
rc=$?
tail -4 gpurun_out/${TAG}_pytest.log
if grep -q "passed" gpurun_out/${TAG}_pytest.log && ! grep -q "failed\|error" gpurun_out/${TAG}_pytest.log; then
  timeout 300 bash scripts/gpu_r2_k.sh ${TAG}b $2
fi
