#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fused_gpu.py tests/test_ops_gpu.py tests/test_infer_gpu.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2i_pytest.log
tail -30 gpurun_out/r2i_pytest.log
python scripts/bench_ops.py > gpurun_out/r2i_ops.jsonl 2> gpurun_out/r2i_ops.err
grep -E "logsumexp|ESS|moments|backward" gpurun_out/r2i_ops.jsonl | cut -c1-200
tail -3 gpurun_out/r2i_ops.err
