// Dev microbenchmarks (B200): latency / throughput of the primitives the env-step kernel leans on.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__global__ void k_lat(long long* out, float* sink, int mode) {
    __shared__ float sm[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (float)((i * 7 + 1) % 1024);
    __syncthreads();
    float v = lane * 1.0f + 1.0f, a0 = v, a1 = v + 1, a2 = v + 2, a3 = v + 3, a4 = v + 4, a5 = v + 5, a6 = v + 6, a7 = v + 7;
    int idx = lane;
    const int N = 256;
    long long t0 = clock64();
    if (mode == 0) {  // dependent SHFL.IDX chain (uniform source)
#pragma unroll 16
        for (int i = 0; i < N; i++) v = __shfl_sync(FULL, v, (i & 7));
    } else if (mode == 1) {  // 8 independent SHFL chains
#pragma unroll 4
        for (int i = 0; i < N / 8; i++) {
            a0 = __shfl_sync(FULL, a0, 1); a1 = __shfl_sync(FULL, a1, 2); a2 = __shfl_sync(FULL, a2, 3); a3 = __shfl_sync(FULL, a3, 4);
            a4 = __shfl_sync(FULL, a4, 5); a5 = __shfl_sync(FULL, a5, 6); a6 = __shfl_sync(FULL, a6, 7); a7 = __shfl_sync(FULL, a7, 8);
        }
        v = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else if (mode == 2) {  // dependent LDS chain
#pragma unroll 16
        for (int i = 0; i < N; i++) idx = (int)sm[idx & 1023];
        v = idx;
    } else if (mode == 3) {  // dependent FFMA chain
#pragma unroll 16
        for (int i = 0; i < N; i++) v = fmaf(v, 1.0000001f, 0.5f);
    } else if (mode == 4) {  // dependent MUFU.RCP chain
#pragma unroll 16
        for (int i = 0; i < N; i++) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); v = r; }
    } else if (mode == 5) {  // SHFL -> FMUL -> FFMA dependent triple chain
#pragma unroll 16
        for (int i = 0; i < N; i++) { float c = __shfl_sync(FULL, v, (i & 7)) * 0.999f; v = fmaf(c, 1.0001f, v); }
    } else if (mode == 6) {  // 8 independent LDS (broadcast address) per iteration
#pragma unroll 4
        for (int i = 0; i < N / 8; i++) {
            a0 += sm[(i * 8 + 0) & 1023]; a1 += sm[(i * 8 + 1) & 1023]; a2 += sm[(i * 8 + 2) & 1023]; a3 += sm[(i * 8 + 3) & 1023];
            a4 += sm[(i * 8 + 4) & 1023]; a5 += sm[(i * 8 + 5) & 1023]; a6 += sm[(i * 8 + 6) & 1023]; a7 += sm[(i * 8 + 7) & 1023];
        }
        v = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else if (mode == 7) {  // STS -> syncwarp -> LDS (other lane) dependent chain
#pragma unroll 8
        for (int i = 0; i < N; i++) { sm[lane + 32 * (threadIdx.x >> 5)] = v; __syncwarp(); v = sm[((lane + 1) & 31) + 32 * (threadIdx.x >> 5)] + 1.f; __syncwarp(); }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
int main() {
    long long* d; float* s;
    cudaMalloc(&d, 148 * 32 * sizeof(long long)); cudaMalloc(&s, 148 * 1024 * sizeof(float));
    const char* names[] = {"dep SHFL chain", "8 indep SHFL chains", "dep LDS chain", "dep FFMA chain", "dep MUFU.RCP chain", "SHFL->FMUL->FFMA chain", "8 indep LDS", "STS->sync->LDS->sync chain"};
    for (int mode = 0; mode < 8; mode++)
        for (int warps : {1, 4, 16}) {
            k_lat<<<148, 32 * warps>>>(d, s, mode);
            cudaDeviceSynchronize();
            k_lat<<<148, 32 * warps>>>(d, s, mode);
            cudaDeviceSynchronize();
            long long h[32];
            cudaMemcpy(h, d, warps * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < warps; i++) avg += h[i];
            printf("%-28s warps/SM %2d: %.1f cycles per op (per warp)\n", names[mode], warps, avg / warps / 256.0);
        }
    return 0;
}
