// Probe: does programmatic dependent launch shorten a captured chain of small dependent kernels on B200?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
template <bool PDL>
__global__ void step_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int work) {
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = in[i];
        for (int k = 0; k < work; ++k) v = fmaf(v, 1.0001f, 0.5f);
        out[i] = v;
    }
}
template <bool PDL>
static void launch(const float* in, float* out, int n, int work, int grid, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, step_kernel<PDL>, in, out, n, work);
}
template <bool PDL>
static float run(int n, int work, int grid, int chain, cudaStream_t s, float* a, float* b) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < chain; ++i) launch<PDL>(i & 1 ? b : a, i & 1 ? a : b, n, work, grid, s);
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return -1; }
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
    cudaEventRecord(e0, s);
    for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return ms / 10 / chain * 1e3f;
}
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    const int nmax = 1 << 24;
    float *a, *b; cudaMalloc(&a, nmax * 4); cudaMalloc(&b, nmax * 4);
    cudaMemset(a, 0, nmax * 4); cudaMemset(b, 0, nmax * 4);
    const int chain = 300;
    for (int n : {1 << 14, 1 << 18, 1 << 21, 1 << 23}) {
        for (int grid : {148, 148 * 8}) {
            float t0 = run<false>(n, 8, grid, chain, s, a, b);
            float t1 = run<true>(n, 8, grid, chain, s, a, b);
            printf("n=%8d grid=%5d  us/kernel: plain %.2f  pdl %.2f\n", n, grid, t0, t1);
        }
    }
    // correctness: the chain result must be identical
    return 0;
}
