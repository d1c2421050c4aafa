#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_vector_gpu.py tests/test_models_gpu.py -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/r2l_pytest.log
tail -40 gpurun_out/r2l_pytest.log
timeout 600 python scripts/bench_configs.py > gpurun_out/r2l_configs.jsonl 2> gpurun_out/r2l_configs.err
cat gpurun_out/r2l_configs.jsonl; tail -3 gpurun_out/r2l_configs.err
