// Timeline of one softmax warp per group inside the attention kernel (clock64 at phase boundaries), common time base.
#define ATT_TRACE 1
#include "../../neurips2024-covomix_b200/csrc/common.cuh"
using namespace covo;
template <int MASK>
void go(AttnArgs a, int stagger) {
    a.stagger = stagger;
    for (int r = 0; r < 2; ++r) attention_tc_kernel<MASK><<<148, ATT_THREADS, ATT_SMEM_BYTES>>>(a);
    printf("mask=0x%x stagger=%d sync: %s\n", MASK, stagger, cudaGetErrorString(cudaDeviceSynchronize()));
    static long long h[4096];
    cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
    const char* nm[7] = {"start", "s_full", "S ld", "max", "exp", "o_full", "P st"};
    long long t00 = h[4 * 256];
    for (int i = 0; i < 4; ++i)
        for (int warp : {4, 8}) {
            long long* e = h + warp * 256 + i * 16;
            printf("  g%d tile %d: start@%6lld |", (warp - 4) / 4, 40 + i, e[0] - t00);
            for (int k = 1; k < 7; ++k) printf(" %s +%lld |", nm[k], e[k] - e[k - 1]);
            printf(" end@%6lld\n", e[6] - t00);
        }
}
int main() {
    const int Bt = 16, N = 1650, H = 16, inner = H * 64;
    __nv_bfloat16 *qkv, *out;
    cudaMalloc(&qkv, (size_t)Bt * N * 3 * inner * 2); cudaMalloc(&out, (size_t)Bt * N * inner * 2);
    cudaMemset(qkv, 0x3c, (size_t)Bt * N * 3 * inner * 2);
    AttnArgs a;
    if (attn_build_args(a, qkv, out, Bt, N, H)) { printf("args: %s\n", err_slot().c_str()); return 1; }
    attn_set_attrs();
    go<0x88>(a, 0);
    go<0x88>(a, 128);
    go<0x92>(a, 128);
    return 0;
}
