// Measures MUFU.EX2 / FFMA / F2FP issue throughput per SM sub-partition on the device (cycles per warp-instruction).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            else if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            else if (MODE == 2) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(a[i])); a[i] = __uint_as_float(r & 0x3fffffff); }
            else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[(i + 4) & 7])); }
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
    const char* names[4] = {"MUFU.EX2", "FFMA", "F2FP.BF16x2", "EX2+FFMA pairs"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
            if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
            double per_smsp_instr = (double)iters * 8 * (mode == 3 ? 2 : 1) * (warps / 4.0);   // warp-instrs per SMSP
            printf("%-16s warps/SM=%2d: %.2f cycles per warp-instruction per SMSP\n", names[mode], warps, c / per_smsp_instr);
        }
    return 0;
}
