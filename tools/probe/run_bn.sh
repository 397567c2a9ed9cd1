python -m pytest tests/test_bf16_kernels_gpu.py tests/test_kernels_gpu.py tests/test_conv2d_gpu.py -m gpu -q -k "bn" > gpurun_out/r2_bn_pytest.log 2>&1
python tools/microbench.py bn 5 > gpurun_out/r2_bn_micro.log 2>&1
