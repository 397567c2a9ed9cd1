python -m pytest tests/test_bf16_kernels_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "depthwise or dwconv or dw" > gpurun_out/r2_dw_pytest.log 2>&1
python tools/microbench.py dw 5 > gpurun_out/r2_dw_micro.log 2>&1
