// Debug build of the tcgen05 GEMM with clock64 stamps (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DLR_TRACE -o build/gemm_trace scratch/gemm_trace.cu -lcuda
#include <vector>
#include <cstdlib>
long long* g_lr_trace = nullptr;
#include "../multimodal_lipread_b200/csrc/abi.cu"
#include "../multimodal_lipread_b200/csrc/gemm_tc.cu"

int main(int argc, char** argv) {
    const int M = atoi(argv[1]), N = atoi(argv[2]), K = atoi(argv[3]);
    const int with_stats = argc > 4 ? atoi(argv[4]) : 0, with_bias = argc > 5 ? atoi(argv[5]) : 0;
    float *A, *B, *C, *bias; double* stats;
    cudaMalloc(&A, (size_t)M * K * 4); cudaMalloc(&B, (size_t)N * K * 4); cudaMalloc(&C, (size_t)M * N * 4);
    cudaMalloc(&bias, N * 4); cudaMalloc(&stats, (64 * N + 2) * 8);
    cudaMemset(A, 0, (size_t)M * K * 4); cudaMemset(B, 0, (size_t)N * K * 4); cudaMemset(bias, 0, N * 4); cudaMemset(stats, 0, (64 * N + 2) * 8);
    const int gx = (M + 127) / 128;
    cudaMalloc(&g_lr_trace, (size_t)gx * 32 * 8);
    cudaMemset(g_lr_trace, 0, (size_t)gx * 32 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaMemset(stats, 0, (64 * N + 2) * 8);
        cudaEventRecord(e0);
        int rc = lr_gemm_tf32(A, K, 0, B, K, 0, C, N, M, N, K, with_bias ? bias : nullptr, with_bias ? 1 : 0, nullptr, 0,
                              with_stats ? stats : nullptr, 1, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rc %d  %.1f us (%s)\n", rc, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<long long> h((size_t)gx * 32);
    cudaMemcpy(h.data(), g_lr_trace, h.size() * 8, cudaMemcpyDeviceToHost);
    const char* names[] = {"start", "setup", "tma0", "tma1", "tma2", "tma3", "tma4", "tma5", "full0", "full1", "full2", "full3", "full4",
                           "full5", "mma_done_issue", "epi_start", "epi_loop_done", "epi_done", "all_done", "box1_begin", "box1_tmem", "box1_smem", "box1_fence", "box1_tma", "box1_stats"};
    for (int b : {0, gx / 2, gx - 1}) {
        printf("CTA %d:", b);
        for (int i = 0; i < 25; ++i) if (h[(size_t)b * 32 + i]) printf(" %s=%lld", names[i], h[(size_t)b * 32 + i] - h[(size_t)b * 32]);
        printf("\n");
    }
    return 0;
}
