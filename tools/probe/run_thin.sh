python -m pytest tests/test_gemm_tc_gpu.py tests/test_bf16_kernels_gpu.py -m gpu -q > gpurun_out/r2_gemm_thin_pytest.log 2>&1
echo "== thin on" > gpurun_out/r2_gemm_thin_micro.log; python tools/microbench.py gemm 5 >> gpurun_out/r2_gemm_thin_micro.log 2>&1
echo "== thin off" >> gpurun_out/r2_gemm_thin_micro.log; LIPREAD_GEMM_THIN=0 python tools/microbench.py gemm 5 >> gpurun_out/r2_gemm_thin_micro.log 2>&1
for st in 2 4; do echo "== thin on, stages $st" >> gpurun_out/r2_gemm_thin_micro.log; LIPREAD_GEMM_STAGES=$st python tools/microbench.py gemm 5 >> gpurun_out/r2_gemm_thin_micro.log 2>&1; done
