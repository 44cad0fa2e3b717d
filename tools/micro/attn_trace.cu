// Timeline of one softmax warp per group inside the attention kernel (clock64 at phase boundaries).
#define ATT_TRACE 1
#include "../../neurips2024-covomix_b200/csrc/common.cuh"
using namespace covo;
int main() {
    const int Bt = 16, N = 1650, H = 16, inner = H * 64;
    __nv_bfloat16 *qkv, *out;
    cudaMalloc(&qkv, (size_t)Bt * N * 3 * inner * 2); cudaMalloc(&out, (size_t)Bt * N * inner * 2);
    cudaMemset(qkv, 0x3c, (size_t)Bt * N * 3 * inner * 2);
    AttnArgs a;
    uint64_t dims[3] = {(uint64_t)3 * inner, (uint64_t)N, (uint64_t)Bt};
    uint64_t str[2] = {(uint64_t)3 * inner * 2, (uint64_t)3 * inner * 2 * N};
    uint32_t box[3] = {64, 128, 1};
    if (make_tmap(&a.tmQKV, qkv, 3, dims, str, box, 0)) { printf("tmap: %s\n", err_slot().c_str()); return 1; }
    a.out = out; a.N = N; a.heads = H; a.inner = inner; a.scale_log2e = 1.4426950408889634f * 0.125f;
    a.n_qt = (N + ATT_BM - 1) / ATT_BM; a.n_items = a.n_qt * H * Bt;
    cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
    for (int r = 0; r < 2; ++r) attention_tc_kernel<<<148, ATT_THREADS, ATT_SMEM_BYTES>>>(a);
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    static long long h[4096];
    cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
    const char* nm[8] = {"iter start", "s_full ok", "S loaded", "max done", "exp+pack done", "p_free ok", "P stored", "O accumulated"};
    for (int warp : {4, 8}) {
        printf("warp %d (group %d)\n", warp, (warp - 4) / 4);
        long long t00 = h[warp * 256];
        for (int i = 0; i < 4; ++i) {
            long long* e = h + warp * 256 + i * 16;
            printf("  it %d: start@%6lld |", 40 + i, e[0] - t00);
            for (int k = 1; k < 8; ++k) printf(" %s +%lld |", nm[k], e[k] - e[k - 1]);
            if (i < 3) printf(" -> next start +%lld", e[16] - e[7]);
            printf("\n");
        }
    }
    return 0;
}
