// Register-level model of the attention kernel's per-tile softmax work: one thread = one query row, 128 scores in
// registers -> row max -> p = exp2(s*c - m*c) -> bf16 pairs (+ row sum).  Measures cycles per 128-score tile per warp for
// several instruction mixes, with 1 and 2 such warps per SM sub-partition (the two query groups of the attention CTA
// sharing one MUFU).  No memory traffic: this is the floor of the exp phase, to be compared with the MUFU bound of
// 128 x 8 = 1024 cycles per warp-tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_loop softmax_loop.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ uint32_t pack_rn(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_trunc(float a, float b) {   // high halves: one PRMT
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
    return r;
}
// 2^x for a pair on the FMA pipe: Cody-Waite split with the 1.5*2^23 trick + degree-3 minimax on [-0.5, 0.5]
__device__ __forceinline__ void exp2_poly2(float& p0, float& p1, float x0, float x1) {
    constexpr float MAGIC = 12582912.f;
    x0 = fmaxf(x0, -125.f);
    x1 = fmaxf(x1, -125.f);
    float t0, t1, n0, n1, f0, f1;
    add2(t0, t1, x0, x1, MAGIC, MAGIC);
    add2(n0, n1, t0, t1, -MAGIC, -MAGIC);
    fma2(f0, f1, n0, n1, -1.f, -1.f, x0, x1);
    float q0, q1;
    fma2(q0, q1, f0, f1, 0.0555041086f, 0.0555041086f, 0.2402264923f, 0.2402264923f);
    fma2(q0, q1, q0, q1, f0, f1, 0.6931471825f, 0.6931471825f);
    fma2(q0, q1, q0, q1, f0, f1, 1.0f, 1.0f);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// MODE: 0 current kernel (scalar FFMA, 4 FADD chains, cvt.rn pack)   1 FFMA2 + FADD2 + cvt.rn
//       2 FFMA2 + cvt.rn, no row sum (sum taken by the tensor core)   3 FFMA2 + FADD2 + truncating PRMT pack
//       4 as 1, every 4th pair through the FMA-pipe polynomial        5 as 1, every 2nd pair polynomial
//       6 FFMA2 + cvt.rn only (no MUFU: what is left next to it)      7 as 2 with 1/4 polynomial   8 as 2 with 1/2 polynomial
//       9 as 3 without the row sum
template <int MODE, bool WITH_MAX>
__global__ void __launch_bounds__(256, 1) k(const float* in, uint32_t* out, long long* cyc, int iters, float c) {
    float s[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 131 + i * 7) & 1023];
    uint32_t chk = 0;
    float m_run = 0.f, l_run = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float m = m_run;
        if (WITH_MAX) {
            float mx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) mx[i] = s[i] + m_run * 1e-6f;     // depends on the previous iteration: not hoistable
#pragma unroll
            for (int i = 8; i + 15 < 128; i += 16)
#pragma unroll
                for (int q = 0; q < 8; ++q) mx[q] = fmax3(mx[q], s[i + 2 * q], s[i + 2 * q + 1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mx[q] = fmax3(mx[q], s[120 + 2 * q], s[121 + 2 * q]);
            m = fmaxf(fmaxf(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5])), fmaxf(mx[6], mx[7]));
        } else {
            m = m_run + 1e-3f;
        }
        const float mc = m * c;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
            float x0, x1, x2, x3, p0, p1, p2, p3;
            if (MODE == 0) {
                x0 = fmaf(s[i], c, -mc); x1 = fmaf(s[i + 1], c, -mc); x2 = fmaf(s[i + 2], c, -mc); x3 = fmaf(s[i + 3], c, -mc);
            } else {
                fma2(x0, x1, s[i], s[i + 1], c, c, -mc, -mc);
                fma2(x2, x3, s[i + 2], s[i + 3], c, c, -mc, -mc);
            }
            const bool poly_a = (MODE == 5 || MODE == 8), poly_b = (MODE == 4 || MODE == 5 || MODE == 7 || MODE == 8) && ((i & 4) || MODE == 5 || MODE == 8);
            if (MODE == 6) { p0 = x0; p1 = x1; p2 = x2; p3 = x3; }
            else {
                if (poly_a && (i & 4)) exp2_poly2(p0, p1, x0, x1); else { p0 = ex2(x0); p1 = ex2(x1); }
                if (poly_b) exp2_poly2(p2, p3, x2, x3); else { p2 = ex2(x2); p3 = ex2(x3); }
            }
            if (MODE == 0) { l0 += p0; l1 += p1; l2 += p2; l3 += p3; }
            else if (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 5) { add2(l0, l1, l0, l1, p0, p1); add2(l2, l3, l2, l3, p2, p3); }
            uint32_t h01, h23;
            if (MODE == 3 || MODE == 9) { h01 = pack_trunc(p0, p1); h23 = pack_trunc(p2, p3); }
            else { h01 = pack_rn(p0, p1); h23 = pack_rn(p2, p3); }
            chk ^= h01 + (h23 << 1);       // stands in for the tcgen05.st of P (2 ALU ops per 4 elements of overhead)
        }
        l_run += (l0 + l1) + (l2 + l3);
        m_run = m * 0.999f + __uint_as_float(chk & 0x3f) * 1e-30f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = chk + __float_as_uint(l_run) + __float_as_uint(m_run);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, bool WITH_MAX>
void run(const char* name, const float* in, uint32_t* out, long long* cyc) {
    const int iters = 512;
    for (int warps : {4, 8}) {
        k<MODE, WITH_MAX><<<148, warps * 32>>>(in, out, cyc, iters, 0.18033688f);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double cy = 0;
        for (int i = 0; i < 148; ++i) cy += h[i];
        cy /= 148;
        printf("%-58s max=%d warps/SMSP=%d: %7.1f cycles per 128-score tile per warp (%7.1f per SMSP per tile round) [%s]\n", name, (int)WITH_MAX,
               warps / 4, cy / iters, cy / iters, cudaGetErrorString(e));
    }
}

int main() {
    float* in; uint32_t* out; long long* cyc;
    cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = -3.f + 6.f * ((i * 2654435761u) >> 8 & 0xffff) / 65536.f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0, false>("0 scalar FFMA + 4 FADD + cvt.rn (round-1 kernel)", in, out, cyc);
    run<0, true>("0 scalar FFMA + 4 FADD + cvt.rn (round-1 kernel)", in, out, cyc);
    run<1, false>("1 FFMA2 + FADD2 + cvt.rn", in, out, cyc);
    run<1, true>("1 FFMA2 + FADD2 + cvt.rn", in, out, cyc);
    run<2, false>("2 FFMA2 + cvt.rn, no row sum", in, out, cyc);
    run<2, true>("2 FFMA2 + cvt.rn, no row sum", in, out, cyc);
    run<3, false>("3 FFMA2 + FADD2 + truncating PRMT pack", in, out, cyc);
    run<9, false>("9 FFMA2 + truncating PRMT pack, no row sum", in, out, cyc);
    run<4, false>("4 FFMA2 + FADD2 + cvt.rn, 1/4 polynomial exp2", in, out, cyc);
    run<4, true>("4 FFMA2 + FADD2 + cvt.rn, 1/4 polynomial exp2", in, out, cyc);
    run<5, false>("5 FFMA2 + FADD2 + cvt.rn, 1/2 polynomial exp2", in, out, cyc);
    run<7, false>("7 FFMA2 + cvt.rn, no row sum, 1/4 polynomial exp2", in, out, cyc);
    run<8, false>("8 FFMA2 + cvt.rn, no row sum, 1/2 polynomial exp2", in, out, cyc);
    run<6, false>("6 FFMA2 + cvt.rn only (no exp)", in, out, cyc);
    return 0;
}
