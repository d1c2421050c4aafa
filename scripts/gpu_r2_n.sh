#!/bin/bash
# 2 GPUs: the NCCL test (train(), graph-captured data-parallel step) and the graphed sampler tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n_gpus.txt
timeout 240 python -m pytest tests/test_multi_gpu.py tests/test_graphed_infer_gpu.py -m gpu -q -rs -v -x 2>&1 | tail -40 > gpurun_out/r2n_pytest.log
cat gpurun_out/r2n_pytest.log | tail -30
