// Issue throughput of the instruction classes that matter in the softmax loop, alone and paired with MUFU.EX2.
#include <cstdio>
#include <cuda_runtime.h>
#define EX2(x) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x))
#define F2FP(x) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(x)); x = __uint_as_float(r & 0x3fffffff); }
#define FADD(x) asm volatile("add.f32 %0, %0, %0;" : "+f"(x))
#define FMX3(x) asm volatile("max.f32 %0, %0, %0, %0;" : "+f"(x))
#define PRMT(x) { unsigned r = __float_as_uint(x); asm volatile("prmt.b32 %0, %0, %0, 0x7632;" : "+r"(r)); x = __uint_as_float(r); }
#define LOP(x) { unsigned r = __float_as_uint(x); asm volatile("and.b32 %0, %0, 0x3fff0000;" : "+r"(r)); x = __uint_as_float(r); }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i + 1);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { EX2(a[i]); F2FP(a[(i + 4) & 7]); }
            if (MODE == 1) { FADD(a[i]); }
            if (MODE == 2) { FMX3(a[i]); }
            if (MODE == 3) { PRMT(a[i]); }
            if (MODE == 4) { LOP(a[i]); }
            if (MODE == 5) { EX2(a[i]); PRMT(a[(i + 4) & 7]); LOP(a[(i + 2) & 7]); }
            if (MODE == 6) { F2FP(a[i]); FADD(a[(i + 4) & 7]); }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    const char* names[7] = {"EX2+F2FP pair", "FADD", "FMNMX3", "PRMT", "LOP3", "EX2+PRMT+LOP3 triple", "F2FP+FADD pair"};
    for (int mode = 0; mode < 7; ++mode) {
        const int warps = 8;
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 4) k<4><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 5) k<5><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 6) k<6><<<148, warps * 32>>>(out, cyc, iters);
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        printf("%-22s: %.2f cycles per group (one of each) per SMSP\n", names[mode], c / ((double)iters * 8 * (warps / 4.0)));
    }
    return 0;
}
