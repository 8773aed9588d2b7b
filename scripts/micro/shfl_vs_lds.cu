// Do warp shuffles share the LSU data pipe with shared-memory loads on sm_100a?  Three kernels, same loop count:
// 8 conflict-free LDS per iteration, 8 SHFL per iteration, and both interleaved.  If T(both) ~ T(lds) + T(shfl) they share a pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o shfl_vs_lds shfl_vs_lds.cu && ./shfl_vs_lds
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    __shared__ float s[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) s[i] = (float)i;
    __syncthreads();
    float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    float v = (float)threadIdx.x;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE & 1) acc0 += s[(idx + 32 * u + it) & 2047];
            if (MODE & 2) { v = __shfl_down_sync(0xffffffffu, v, 1); acc1 += v; }
        }
        acc2 += acc0 * 1.0001f; acc3 += acc1;
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000, grid = 148 * 8;
    float ms[4] = {};
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a); k<1><<<grid, 256>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms[1], a, b);
        cudaEventRecord(a); k<2><<<grid, 256>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms[2], a, b);
        cudaEventRecord(a); k<3><<<grid, 256>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms[3], a, b);
    }
    // per SM: 8 CTAs x 8 warps x iters x 8 ops
    const double ops = 8.0 * 8 * iters * 8;
    printf("LDS only %.3f ms (%.2f cyc/warp-op/SM at 1.965 GHz) | SHFL only %.3f ms (%.2f) | both %.3f ms (sum %.3f, max %.3f)\n", ms[1],
           ms[1] * 1.965e6 / ops, ms[2], ms[2] * 1.965e6 / ops, ms[3], ms[1] + ms[2], ms[1] > ms[2] ? ms[1] : ms[2]);
    return 0;
}
