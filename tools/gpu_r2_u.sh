#!/bin/bash
# Round-2 GPU session U (1 GPU, < 40 s): the 33..64-tap stem wgrad tests and the heads tests.
mkdir -p gpurun_out
IIC_RUN_UNVALIDATED=1 timeout 35 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=line -p no:cacheprovider -k "stem_wgrad_on_tensor or test_heads" > gpurun_out/u_tests.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/u_tests.log
