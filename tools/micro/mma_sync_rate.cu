// Microbenchmark: sustained rate of the legacy tensor path (mma.sync.m16n8k16 f16 x f16 -> f32, SASS HMMA) on sm_100a.
// DESIGN.md §7-1 asks what a banded-Toeplitz formulation of K1's 11-tap passes would have to run at; this measures the ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) hmma_loop(float *out, int iters) {
    unsigned a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x38003800u};
    float acc[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)  // eight independent accumulators per warp
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[k][0]), "+f"(acc[k][1]), "+f"(acc[k][2]), "+f"(acc[k][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        for (int j = 0; j < 4; ++j) s += acc[k][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000;
    for (int bps : {1, 2, 4, 8}) {
        hmma_loop<<<sms * bps, 256>>>(out, 100);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        hmma_loop<<<sms * bps, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double mmas = (double)sms * bps * 8 /*warps*/ * 8 * iters;
        printf("mma.sync m16n8k16 f16->f32: %d CTAs/SM x 8 warps: %.1f TFLOP/s (%.2f MMA per SM per clock at 1.965 GHz)\n", bps,
               mmas * 4096.0 / (ms * 1e-3) / 1e12, mmas / sms / (ms * 1e-3 * 1.965e9));
    }
    return 0;
}
