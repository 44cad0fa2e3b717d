// Stand-alone check + timing of attention_tc_kernel variants (exponential mix x group stagger) against the naive
// CUDA-core kernel, at the C3 (Bt = 16, N = 1650) and C2 (Bt = 2, N = 650) shapes and a few ragged ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o attn_bench attn_bench.cu -lcuda
//   ./attn_bench [quick]
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../neurips2024-covomix_b200/csrc/common.cuh"
using namespace covo;

static uint32_t rng_state = 12345u;
static float frand() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xffffff) / 16777216.0f;
}
static float nrand() { return sqrtf(-2.f * logf(frand() + 1e-7f)) * cosf(6.2831853f * frand()); }

template <int MASK, int HO>
static float run(const AttnArgs& a, int reps) {
    const int grid = a.n_items < 148 ? a.n_items : 148;
    cudaFuncSetAttribute(attention_tc_kernel<MASK, HO>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
    for (int r = 0; r < 2; ++r) attention_tc_kernel<MASK, HO><<<grid, ATT_THREADS, ATT_SMEM_BYTES>>>(a);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) attention_tc_kernel<MASK, HO><<<grid, ATT_THREADS, ATT_SMEM_BYTES>>>(a);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        exit(2);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1;
    if (attn_set_attrs() != COVO_OK) { printf("attrs: %s\n", err_slot().c_str()); return 1; }
    struct Shape { int Bt, N, H; float qscale; };
    std::vector<Shape> shapes = {{2, 1, 1, 1.f}, {1, 127, 2, 1.f}, {2, 128, 1, 1.f}, {1, 129, 2, 1.f}, {1, 300, 2, 6.f}, {2, 650, 16, 1.f},
                                 {2, 650, 16, 6.f}, {16, 1650, 16, 1.f}, {16, 1650, 16, 6.f}};
    for (const Shape& sh : shapes) {
        const int Bt = sh.Bt, N = sh.N, H = sh.H, inner = H * 64;
        const size_t nq = (size_t)Bt * N * 3 * inner, no = (size_t)Bt * N * inner;
        std::vector<__nv_bfloat16> hq(nq);
        for (size_t i = 0; i < nq; ++i) {
            const int col = i % (3 * inner);
            float v = nrand();
            if (col < inner) v *= sh.qscale;                 // peaky scores: running max moves by > 2^8 in the first tiles
            hq[i] = __float2bfloat16(v);
        }
        __nv_bfloat16 *qkv, *out, *ref;
        cudaMalloc(&qkv, nq * 2);
        cudaMalloc(&out, no * 2);
        cudaMalloc(&ref, no * 2);
        cudaMemcpy(qkv, hq.data(), nq * 2, cudaMemcpyHostToDevice);
        dim3 g((N + 7) / 8, H, Bt);
        naive_attention_kernel<<<g, 256>>>(qkv, ref, N, H, inner, 0.125f);
        cudaDeviceSynchronize();
        std::vector<__nv_bfloat16> href(no), hout(no);
        cudaMemcpy(href.data(), ref, no * 2, cudaMemcpyDeviceToHost);
        AttnArgs a;
        if (attn_build_args(a, qkv, out, Bt, N, H) != COVO_OK) { printf("args: %s\n", err_slot().c_str()); return 1; }
        const double flops = 4.0 * N * (double)N * 64 * H * Bt;
        const bool big = (size_t)Bt * N >= 1300;
        for (int poly = 0; poly < 3; ++poly) {
            for (int stagger : {-1, 0, 64, 96, 128}) {    // -1: token code compiled out (HANDOFF = 0); 0: free running; else the token hand-off point (of 128 scores)
                if (!big && stagger != 0 && stagger != 128 && stagger != -1) continue;
                if (quick && poly == 0 && stagger != 0 && stagger != 128) continue;
                a.stagger = stagger < 0 ? 0 : stagger;
                cudaMemset(out, 0xff, no * 2);
                const int reps = big ? 10 : 1;
                float ms;
                if (stagger == -1) ms = poly == 0 ? run<0, 0>(a, reps) : (poly == 1 ? run<0x88, 0>(a, reps) : run<0x92, 0>(a, reps));
                else if (stagger == 64) ms = poly == 0 ? run<0, 64>(a, reps) : (poly == 1 ? run<0x88, 64>(a, reps) : run<0x92, 64>(a, reps));
                else if (stagger == 96) ms = poly == 0 ? run<0, 96>(a, reps) : (poly == 1 ? run<0x88, 96>(a, reps) : run<0x92, 96>(a, reps));
                else ms = poly == 0 ? run<0, 128>(a, reps) : (poly == 1 ? run<0x88, 128>(a, reps) : run<0x92, 128>(a, reps));
                cudaMemcpy(hout.data(), out, no * 2, cudaMemcpyDeviceToHost);
                double num = 0, den = 0, mx = 0;
                int bad = 0;
                for (size_t i = 0; i < no; ++i) {
                    const double x = __bfloat162float(hout[i]), r = __bfloat162float(href[i]);
                    if (!std::isfinite(x)) ++bad;
                    num += (x - r) * (x - r);
                    den += r * r;
                    if (fabs(x - r) > mx) mx = fabs(x - r);
                }
                printf("Bt=%2d N=%4d H=%2d qscale=%.0f poly=%d stagger=%4d: %8.1f us %7.1f TFLOP/s  rel-L2 %.2e max-abs %.2e nonfinite %d %s\n", Bt, N, H,
                       sh.qscale, poly, stagger, ms * 1e3, flops / ms / 1e9, sqrt(num / (den + 1e-30)), mx, bad,
                       (bad || sqrt(num / (den + 1e-30)) > 5e-3) ? "FAIL" : "ok");
                fflush(stdout);
            }
        }
        cudaFree(qkv);
        cudaFree(out);
        cudaFree(ref);
    }
    return 0;
}
