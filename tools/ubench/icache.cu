// Dev microbenchmark: instruction-cache capacity seen by the warps of one SM (B200).
// B straight-line code blocks of ~1 KB each (64 FFMA with unique immediates) are executed round-robin by 16 warps per CTA,
// one CTA per SM, either all warps at the same block (sync) or each warp starting at its own offset (desync).
// Output: cycles per warp-instruction vs footprint.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ID>
__device__ __noinline__ void blk(float& a, float& b, float& c, float& d) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
        a = fmaf(a, 1.0f + 1e-6f * (ID * 64 + 4 * k + 1), 1e-3f * (ID + k));
        b = fmaf(b, 1.0f + 1e-6f * (ID * 64 + 4 * k + 2), 2e-3f * (ID + k));
        c = fmaf(c, 1.0f + 1e-6f * (ID * 64 + 4 * k + 3), 3e-3f * (ID + k));
        d = fmaf(d, 1.0f + 1e-6f * (ID * 64 + 4 * k + 4), 4e-3f * (ID + k));
    }
}
template <int I, int B>
struct Call {
    static __device__ __forceinline__ void run(int id, float& a, float& b, float& c, float& d) {
        if constexpr (I < B) {
            if (id == I) { blk<I>(a, b, c, d); return; }
            Call<I + 1, B>::run(id, a, b, c, d);
        }
    }
};
// binary dispatch keeps the dispatch cost O(log B)
template <int LO, int HI>
struct Disp {
    static __device__ __forceinline__ void run(int id, float& a, float& b, float& c, float& d) {
        if constexpr (HI - LO == 1) blk<LO>(a, b, c, d);
        else {
            constexpr int MID = (LO + HI) / 2;
            if (id < MID) Disp<LO, MID>::run(id, a, b, c, d); else Disp<MID, HI>::run(id, a, b, c, d);
        }
    }
};
template <int B>
__global__ void __launch_bounds__(512) k(int iters, int desync, float* out, long long* cyc) {
    float a = threadIdx.x, b = 1.f, c = 2.f, d = 3.f;
    const int warp = threadIdx.x >> 5;
    int id = desync ? (warp * B / 16) % B : 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters * B; i++) {
        Disp<0, B>::run(id, a, b, c, d);
        id = id + 1 == B ? 0 : id + 1;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
    if (a + b + c + d == 12345.f) out[0] = a;
}
template <int B>
void run(int warps) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    for (int desync = 0; desync < 2; desync++) {
        const int iters = 4096 / B > 4 ? 4096 / B : 4;
        k<B><<<148, 32 * warps>>>(2, desync, out, cyc);  // warm
        cudaMemset(cyc, 0, 8);
        k<B><<<148, 32 * warps>>>(iters, desync, out, cyc);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per_warp = (double)h / (148.0 * warps), inst = (double)iters * B * 68.0;
        printf("footprint %4d KB  warps %2d  %s  cycles/warp-inst %.3f  SM inst/cycle %.3f\n", B, warps, desync ? "desync" : "sync  ",
               per_warp / inst, warps * inst / per_warp);
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int warps : {16, 4}) {
        run<8>(warps); run<16>(warps); run<32>(warps); run<48>(warps); run<64>(warps); run<96>(warps); run<128>(warps); run<192>(warps); run<256>(warps); run<512>(warps);
    }
    return 0;
}
