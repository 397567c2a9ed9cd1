python -m pytest tests/test_gemm_tc_gpu.py tests/test_bf16_kernels_gpu.py -m gpu -q -x > gpurun_out/r2_gemm_mt_pytest.log 2>&1
for mt in 1 2 4 8; do for st in 2 4; do echo "== MT=$mt STAGES=$st" ; LIPREAD_GEMM_MT=$mt LIPREAD_GEMM_STAGES=$st python tools/microbench.py gemm 5; done; done > gpurun_out/r2_gemm_mt_micro.log 2>&1
echo "== auto" >> gpurun_out/r2_gemm_mt_micro.log; python tools/microbench.py gemm 5 >> gpurun_out/r2_gemm_mt_micro.log 2>&1
