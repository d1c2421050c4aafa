#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_infer_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2m_pytest.log
tail -6 gpurun_out/r2m_pytest.log
python scripts/bench_ops.py > gpurun_out/r2m_ops.jsonl 2> gpurun_out/r2m_ops.err
grep -E "logsumexp|ESS|moments|backward" gpurun_out/r2m_ops.jsonl | cut -c1-160
tail -3 gpurun_out/r2m_ops.err
