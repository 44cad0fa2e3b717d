// Does the packed half-precision exponential double the MUFU throughput?  Cycles per warp-instruction per SM
// sub-partition for ex2.approx.ftz.f32, ex2.approx.f16x2, ex2.approx.ftz.bf16x2, cvt.rn.f16x2.f32 and the f16 -> f32 unpack,
// plus the full softmax inner-loop candidates (per PAIR of scores):
//   A (current): 2 FFMA + 2 EX2.f32 + 1 cvt.bf16x2 + 2 FADD
//   B (packed) : 2 FFMA + 1 cvt.f16x2 + 1 EX2.f16x2 + 2 (f16 -> f32) + 2 FADD
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = -0.001f * (threadIdx.x + i + 1);
        u[i] = 0xb800b800u + threadIdx.x + i;       // two small negative halves
    }
    float acc0 = 0.f, acc1 = 0.f;
    const float c = 0.18f, mc = 0.3f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            else if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            else if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
            else if (MODE == 3) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(a[i])); a[i] = __uint_as_float(u[i] & 0x3fffffffu); }
            else if (MODE == 4) {      // loop A
                float x0 = fmaf(a[i], c, -mc), x1 = fmaf(a[(i + 1) & 7], c, -mc), p0, p1;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
                unsigned r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(p1), "f"(p0));
                acc0 += p0;
                acc1 += p1;
                u[i] ^= r;
            } else {                   // loop B
                float x0 = fmaf(a[i], c, -mc), x1 = fmaf(a[(i + 1) & 7], c, -mc);
                unsigned r;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
                const __half2 h = *reinterpret_cast<const __half2*>(&r);
                const float2 f = __half22float2(h);
                acc0 += f.x;
                acc1 += f.y;
                u[i] ^= r;
            }
        }
    }
    long long t1 = clock64();
    float s = acc0 + acc1;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    const char* names[6] = {"EX2.f32", "EX2.f16x2", "EX2.bf16x2", "cvt.f16x2.f32", "softmax pair A (f32 exp)", "softmax pair B (f16x2 exp)"};
    for (int mode = 0; mode < 6; ++mode)
        for (int warps : {8, 16}) {
            if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 4) k<4><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 5) k<5><<<148, warps * 32>>>(out, cyc, iters);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
            double per_smsp = (double)iters * 8 * (warps / 4.0);
            printf("%-28s warps/SM=%2d: %.2f cycles per warp-%s per SMSP  (%s)\n", names[mode], warps, c / per_smsp,
                   mode >= 4 ? "pair-of-scores" : "instruction", cudaGetErrorString(e));
        }
    return 0;
}
