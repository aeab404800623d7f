// Dependent-issue latencies on the box (one warp, one CTA): what a serial per-sample recurrence pays per operation.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/latency.cu -o tools/micro/latency && tools/micro/latency
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k(double* out, long long* cyc, double a, double b, float fa, float fb) {
    double x = a; float f = fa; long long t0, t1;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    t1 = clock64(); cyc[0] = t1 - t0;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + b;
    t1 = clock64(); cyc[1] = t1 - t0;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, fb, fa);
    t1 = clock64(); cyc[2] = t1 - t0;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; ++i) f = __shfl_sync(0xffffffffu, f, (threadIdx.x + 1) & 31);
    t1 = clock64(); cyc[3] = t1 - t0;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; ++i) x = (double) ((float) x) + b;       // f64 -> f32 -> f64 round trip + add
    t1 = clock64(); cyc[4] = t1 - t0;
    __shared__ float sm[64];
    sm[threadIdx.x] = f; sm[threadIdx.x + 32] = f; __syncwarp();
    t0 = clock64();
    int idx = threadIdx.x;
    #pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int) sm[idx & 63] & 31;        // dependent shared-memory loads
    t1 = clock64(); cyc[5] = t1 - t0;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; ++i) x = x / (b + x);                       // dependent IEEE double division
    t1 = clock64(); cyc[6] = t1 - t0;
    out[threadIdx.x] = x + f + idx;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 256 * 8); cudaMalloc(&c, 64);
    for (int warps = 1; warps <= 16; warps *= 4) {
        k<<<1, 32 * warps>>>(o, c, 1.0000001, 0.9999999, 1.0001f, 0.9999f);
        k<<<1, 32 * warps>>>(o, c, 1.0000001, 0.9999999, 1.0001f, 0.9999f);
        long long h[8]; cudaMemcpy(h, c, 56, cudaMemcpyDeviceToHost);
        printf("warps/CTA %2d  cycles per dependent op: DFMA %.1f  DADD %.1f  FFMA %.1f  SHFL %.1f  F64->F32->F64+DADD %.1f  LDS(dependent) %.1f  DDIV %.1f\n", warps,
               h[0] / (double) N, h[1] / (double) N, h[2] / (double) N, h[3] / (double) N, h[4] / (double) N, h[5] / (double) N, h[6] / (double) N);
    }
    return 0;
}
