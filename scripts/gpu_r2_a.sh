#!/bin/bash
# round 2, GPU call A (2 GPUs): full GPU test suite incl. the 2-rank NCCL test and the reference's own unittests,
# then bench at N=1 and N=2 with the new legs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
AESMC_REF_TESTS_LONG=0 timeout 900 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -150 > gpurun_out/r2a_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs -v 2>&1 | tail -15 > gpurun_out/r2a_multi_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench_2gpu.json 2> gpurun_out/r2a_bench_2gpu.err
tail -c 1500 gpurun_out/r2a_pytest.log
tail -c 600 gpurun_out/r2a_multi_gpu.log
tail -c 3000 gpurun_out/r2a_bench_2gpu.json
tail -c 500 gpurun_out/r2a_bench_2gpu.err
