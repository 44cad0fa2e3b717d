// Text-to-semantic (CoSingle / CoMix) on sm_100a: TextToSemantic.generate (covomix/covomix_model/text2semantic.py:659-848).
//
// Two parts:
//  * the source side (embedding, non-causal source_transformer, cross-attention k/v of every decoder layer) runs once per
//    call as a handful of fp32 kernels -- a few hundred text tokens, < 1 % of the work;
//  * the autoregressive loop is ONE persistent cooperative kernel (t2s_decode_kernel): all SMs walk the stage list of a
//    decoding step together, separated by grid barriers, and loop over steps without returning to the host -- no kernel
//    launch, no host round trip and no re-read of the prefix per token (the reference re-embeds the whole prefix,
//    re-projects the cross-attention context and launches ~200 kernels per step).  A step is a chain of skinny
//    products (B <= 8 rows against 46 M parameters, 93 MB as bf16) bound by the latency of its 34 dependent stages, not
//    by FLOPs or bytes: with bf16 matrices the products run on warp-level tensor cores (mma.sync, batch = n dimension,
//    K split over the warps of a CTA; t2s_gemv_mma), with fp32 matrices on CUDA cores (t2s_gemv); fp32 accumulation,
//    norms, softmax and logits either way.  Single-query attention over a bf16 (or fp32) KV cache, the sampler
//    (top-k filter + Gumbel arg-max, text2semantic.py:104-132) and the EOS logic (:804-826) are stages of the same kernel.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace covo {

constexpr int T2S_MAX_DEPTH = 8;
constexpr int T2S_THREADS = 512;
constexpr int T2S_WARPS = T2S_THREADS / 32;
constexpr int T2S_DH = 64;            // dim_head (text2semantic.py:432 default; the only value supported)
constexpr int T2S_PART = 66;          // attention partial: m, l, o[64]

// ======================================================================================================================
// source side: small fp32 kernels
// ======================================================================================================================

// ids int64 [B,S1] (EOS already set by the host mirror of set_eos_id, text2semantic.py:57-66) -> x f32 [B*S1, D],
// mask[b,s] = ids != pad (:726-727)
__global__ void t2s_embed_text_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                      float* __restrict__ x, uint8_t* __restrict__ mask, int D, long long pad_id,
                                      long long n_rows) {
    const int row = blockIdx.x;
    long long id = ids[row];
    if (threadIdx.x == 0) mask[row] = id != pad_id;
    if (id < 0 || id >= n_rows) id = 0;
    const float* e = table + static_cast<size_t>(id) * D;
    for (int j = threadIdx.x; j < D; j += blockDim.x) x[static_cast<size_t>(row) * D + j] = e[j];
}

// RMSNorm (text2semantic.py:143-151), fp32 in/out, one warp per row
__global__ void t2s_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, float* __restrict__ out,
                                   int M, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + static_cast<size_t>(row) * D;
    float ss = 0.f;
    for (int j = lane; j < D; j += 32) ss += xr[j] * xr[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = sqrtf(static_cast<float>(D)) / fmaxf(sqrtf(ss), 1e-12f);
    for (int j = lane; j < D; j += 32) out[static_cast<size_t>(row) * D + j] = xr[j] * scale * gamma[j];
}

// interleaved-pair rotary (rotary_embedding_torch.py:36-52) applied in place to `n_cols` leading columns of t [M, ld];
// position = row % S1.  tab: float2 (cos, sin) [pos][32]
__global__ void t2s_rope_rows_kernel(float* __restrict__ t, const float2* __restrict__ tab, int M, int ld, int n_cols,
                                     int S1) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = n_cols / 2;
    if (idx >= M * pairs) return;
    const int row = idx / pairs, p = idx % pairs;
    const int j = p % (T2S_DH / 2);
    const float2 cs = tab[static_cast<size_t>(row % S1) * (T2S_DH / 2) + j];
    float* q = t + static_cast<size_t>(row) * ld + 2 * p;
    const float x0 = q[0], x1 = q[1];
    q[0] = x0 * cs.x - x1 * cs.y;
    q[1] = x1 * cs.x + x0 * cs.y;
}

// Non-causal masked attention of the source transformer (attend_t2s.py:127-171): one CTA per (b, h, query).
// q [B*S1, inner], kv [B*S1, 2*inner] (k | v), out [B*S1, inner].  S1 is a few hundred at most.
__global__ void __launch_bounds__(128) t2s_enc_attention_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                                const uint8_t* __restrict__ mask, float* __restrict__ out,
                                                                int S1, int H) {
    extern __shared__ float sm[];
    float* sc = sm;                 // [S1]
    float* sq = sm + S1;            // [64]
    __shared__ float red[4];
    const int i = blockIdx.x % S1, h = (blockIdx.x / S1) % H, b = blockIdx.x / (S1 * H);
    const int inner = H * T2S_DH;
    const int tid = threadIdx.x;
    if (tid < T2S_DH) sq[tid] = q[(static_cast<size_t>(b) * S1 + i) * inner + h * T2S_DH + tid] * 0.125f;
    __syncthreads();
    float mx = -3.402823466e38f;
    for (int j = tid; j < S1; j += blockDim.x) {
        const float* kr = kv + (static_cast<size_t>(b) * S1 + j) * 2 * inner + h * T2S_DH;
        float s = 0.f;
#pragma unroll 16
        for (int d = 0; d < T2S_DH; ++d) s = fmaf(sq[d], kr[d], s);
        if (!mask[b * S1 + j]) s = -3.402823466e38f;
        sc[j] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < S1; j += blockDim.x) {
        const float p = expf(sc[j] - mx);
        sc[j] = p;
        sum += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    __syncthreads();
    sum = red[0] + red[1] + red[2] + red[3];
    if (tid < T2S_DH) {
        float o = 0.f;
        for (int j = 0; j < S1; ++j)
            o = fmaf(sc[j], kv[(static_cast<size_t>(b) * S1 + j) * 2 * inner + inner + h * T2S_DH + tid], o);
        out[(static_cast<size_t>(b) * S1 + i) * inner + h * T2S_DH + tid] = o / sum;
    }
}

__global__ void t2s_add_kernel(float* __restrict__ x, const float* __restrict__ y, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) x[i] += y[i];
}

// GEGLU (text2semantic.py:155-158): h [M, 2*fi] -> g [M, fi] = gelu(h[:, fi:]) * h[:, :fi]
__global__ void t2s_geglu_kernel(const float* __restrict__ h, float* __restrict__ g, int M, int fi) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<size_t>(M) * fi) return;
    const size_t m = i / fi, j = i % fi;
    g[i] = gelu_erf(h[m * 2 * fi + fi + j]) * h[m * 2 * fi + j];
}

// kv [B*S1, 2*inner] of the encoded text + null_kv [2,H,64] -> cross-attention context of one decoder layer (null first,
// :253-257) in the layout the decode kernel reads (see T2SLayerW), n_alloc = n_ctx rounded up to 32;
// cmask [B][1+S1] = 1 | source_mask (:259-260)
__global__ void t2s_ctx_scatter_kernel(const float* __restrict__ kv, const float* __restrict__ null_kv,
                                       const uint8_t* __restrict__ mask, void* __restrict__ ck, void* __restrict__ cv,
                                       uint8_t* __restrict__ cmask, int B, int S1, int H, int n_alloc, int kv16) {
    const int inner = H * T2S_DH, n_ctx = S1 + 1;
    const size_t n = static_cast<size_t>(B) * H * n_ctx * (T2S_DH / 2);
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dp = i % (T2S_DH / 2);                       // dim pair
    const int j = (i / (T2S_DH / 2)) % n_ctx;
    const int h = (i / (static_cast<size_t>(T2S_DH / 2) * n_ctx)) % H;
    const int b = i / (static_cast<size_t>(T2S_DH / 2) * n_ctx * H);
    float2 k2, v2;
    if (j == 0) {
        k2 = *reinterpret_cast<const float2*>(null_kv + h * T2S_DH + 2 * dp);
        v2 = *reinterpret_cast<const float2*>(null_kv + (H + h) * T2S_DH + 2 * dp);
    } else {
        const float* r = kv + (static_cast<size_t>(b) * S1 + (j - 1)) * 2 * inner + h * T2S_DH + 2 * dp;
        k2 = *reinterpret_cast<const float2*>(r);
        v2 = *reinterpret_cast<const float2*>(r + inner);
    }
    const size_t bh = static_cast<size_t>(b) * H + h;
    if (kv16) {
        const __nv_bfloat162 kp = __floats2bfloat162_rn(k2.x, k2.y), vp = __floats2bfloat162_rn(v2.x, v2.y);
        static_cast<uint32_t*>(ck)[(bh * (n_alloc / 32) + j / 32) * 1024 + dp * 32 + (j & 31)] = *reinterpret_cast<const uint32_t*>(&kp);
        static_cast<uint32_t*>(cv)[(bh * n_alloc + j) * 32 + dp] = *reinterpret_cast<const uint32_t*>(&vp);
    } else {
        *reinterpret_cast<float2*>(static_cast<float*>(ck) + (bh * n_alloc + j) * T2S_DH + 2 * dp) = k2;
        *reinterpret_cast<float2*>(static_cast<float*>(cv) + (bh * n_alloc + j) * T2S_DH + 2 * dp) = v2;
    }
    if (h == 0 && dp == 0) cmask[b * n_ctx + j] = j == 0 ? 1 : mask[b * S1 + j - 1];
}

// ======================================================================================================================
// the persistent decode kernel
// ======================================================================================================================
constexpr int T2S_MAXS = 64;          // max key groups per (b, h)
constexpr int T2S_BLOCKS_PER_WARP = 4; // 32-key blocks a warp walks before a sequence is split across CTAs

struct T2SLayerW {
    const float* sa_gamma;
    const void* sa_qkv;      // [3*inner, Dt]   rows: q | k | v   (to_q.0.weight ; to_kv.0.weight)
    const void* sa_out;      // [Dt, inner]
    const float* ca_gamma;
    const void* ca_q;        // [inner, Dt]
    const void* ca_out;      // [Dt, inner]
    const float* ff_gamma;
    const void* ff1;         // [2*ffi, Dt]     rows: x | gate
    const float* ff1_b;
    const void* ff2;         // [Dt, ffi_pad]
    const float* ff2_b;
    // KV storage, per (b, h), n_alloc = positions rounded up to 32:
    //   fp32 matrices : K, V fp32 [n_alloc][64]
    //   bf16 matrices : K bf16x2 words, blocked-transposed [n_alloc/32][32 dim pairs][32 positions] (a lane owns a key and
    //                   reads one coalesced word per dim pair); V bf16x2 words [n_alloc][32] (a lane owns two dims)
    const void* ctx_k;       // cross-attention context (null kv first), n_alloc = n_ctx_r
    const void* ctx_v;
    void* kcache;            // self-attention cache (rotated keys), n_alloc = max_len_r
    void* vcache;
};

struct T2SDecArgs {
    T2SLayerW L[T2S_MAX_DEPTH];
    int depth, B, Dt, inner, H, ffi, ffi_pad, n_out, demb, n_logits, n_ctx, max_len, n_ctx_r, max_len_r, topk, ignore_eos, dbg_mode;
    float temperature;
    long long eos_id;
    const float* emb;            // [n_logits, demb] fp32: input embedding and (tied) logit projection
    const float* start;          // [Dt]
    const float* final_gamma;
    const float2* rope;          // [max_len][32]
    const uint8_t* ctx_mask;     // [B][n_ctx]
    float* x;                    // [B][Dt]   residual stream of the current position
    float* q;                    // [B][inner]
    float* attn;                 // [B][inner]   merged attention output
    float* part;                 // [B*H][T2S_MAXS][66]
    unsigned* part_cnt;          // [B*H] arrival counters of the split merge
    float* hbuf;                 // [B][ffi_pad]   (padding columns stay zero)
    float* logits;               // [n_out][B][n_logits]
    const float* u;              // [max_len][n_out][B][n_logits]
    const long long* forced;     // [B][n_out][max_len] or null
    long long* tokens;           // [B][n_out][max_len]
    float* logits_out;           // [max_len][n_out][B][n_logits] or null
    int* result;                 // [0] steps done, [1] stopped by EOS, [2] abort flag
    int* eos_flags;              // [n_out][B]
    unsigned* barrier;
};

// All CTAs are co-resident (cooperative launch).  Arrive = one red.release (orders this CTA's earlier writes, made visible to
// thread 0 by the __syncthreads); wait = relaxed polling.  Every buffer another CTA wrote is read with __ldcg (L2), so no
// gpu-scope acquire fence -- which would also throw away L1 -- is needed after the wait.  A stuck barrier sets the abort flag
// instead of hanging the device.
__device__ __forceinline__ bool t2s_grid_barrier(const T2SDecArgs& a, unsigned& epoch) {
    __shared__ int s_abort;
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(a.barrier), "r"(1u) : "memory");
        const long long t0 = clock64();
        int ab = 0;
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.barrier) : "memory");
            if (v < epoch && clock64() - t0 > 4000000000ll) {
                a.result[2] = 1;
                ab = 1;
                break;
            }
        } while (v < epoch);
        s_abort = ab;       // a CTA that gives up leaves the others to their own time-out
    }
    __syncthreads();
    return s_abort != 0;
}

// 16-byte load of a decoder matrix.  The matrices are streamed once per decoding step (92 MB as bf16 -- three quarters of
// L2); marking them evict-first keeps them from flushing the small, latency-critical working set (activations, KV cache,
// cross-attention context) out of L2 every step.
__device__ int t2s_dbg_plain_loads = 0;
__device__ __forceinline__ uint4 t2s_ld_stream(const void* p) {
    uint4 v;
    if (t2s_dbg_plain_loads) return *reinterpret_cast<const uint4*>(p);
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}

template <class WT>
struct WVec;
template <>
struct WVec<__nv_bfloat16> {
    static constexpr int N = 8;     // elements per 16-byte load
    __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&w)[8]) {
        const uint4 v = t2s_ld_stream(p);
        const uint32_t r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            w[2 * i] = __uint_as_float(r[i] << 16);
            w[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
        }
    }
};
template <>
struct WVec<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void load(const float* p, float (&w)[4]) {
        const uint4 v = t2s_ld_stream(p);
        w[0] = __uint_as_float(v.x), w[1] = __uint_as_float(v.y), w[2] = __uint_as_float(v.z), w[3] = __uint_as_float(v.w);
    }
};

// debug tracer (COVO_T2S_TRACE=<step>): CTA 0 stamps clock64 at stage boundaries of one decoding step
__device__ long long t2s_trace_buf[128];
__device__ int t2s_trace_step = -1;
#define T2S_MARK(id)                                                                              \
    do {                                                                                          \
        if (blockIdx.x == 0 && threadIdx.x == 0 && step == t2s_trace_step) t2s_trace_buf[id] = clock64(); \
    } while (0)
__device__ int t2s_dbg_attn_groups = 0;   // debug: force the number of key groups per (b, h) (COVO_T2S_ATTN_GROUPS)
__device__ int t2s_dbg_no_prefetch = 0;
__device__ int t2s_dbg_skip_gemv = 0;     // debug: time the kernel without the matrix products (tools/t2s_bench.py)

template <int N>
__device__ __forceinline__ float t2s_select(const float (&v)[N], int i) {
    float r = v[0];
#pragma unroll
    for (int j = 1; j < N; ++j) r = (i == j) ? v[j] : r;
    return r;
}

// Activation rows live in shared memory.  For 16-bit weights a lane's 16-byte weight load covers 8 consecutive k; the two
// float4 halves of the matching activations are stored `half` = K/2 floats apart so that consecutive lanes read consecutive
// 16-byte words (conflict-free; the linear layout is a 2-way bank conflict on every read, which bounds the B = 8 stages).
// half == 0: linear layout (fp32 weights: 4 k per load).
__device__ __forceinline__ int t2s_sx_index(int j, int half) {
    return half ? ((((j >> 3) << 2) | (j & 3)) + ((j >> 2) & 1) * half) : j;
}
template <class WT>
__device__ __forceinline__ int t2s_half(int K) {
    return WVec<WT>::N == 8 ? K / 2 : 0;
}

// Skinny product: unit u covers R rows (r0 = u * unit_stride, r1 = r0 + pair_off when R == 2) of W [*, ldw] against the NB
// activation rows in shared memory (layout t2s_sx_index with half = t2s_half<WT>(K)).  Units are dealt to the warps of the
// grid CTA-interleaved, so stages with fewer units than warps still pull through every SM.  Lane b < NB first calls
// pre(r0, r1, b) (a prefetch -- e.g. the residual -- issued BEFORE the weight loads so that its latency hides behind them)
// and, once the sums are complete, epi(r0, r1, b, v0, v1, prefetched).  K % WVec::N == 0; W rows are 16-byte aligned.
// Optional grouping (the tied logit projection of the two output streams in one pass): unit u belongs to group
// u / units_per_group, which reads its activations at sx + group * sx_group_stride and reports rows offset by
// group * row_group_stride to pre / epi; the weight rows are those of u % units_per_group.
template <class WT, int NB, int R, class Pre, class Epi>
__device__ __forceinline__ void t2s_gemv(const WT* __restrict__ W, int ldw, int n_units, int unit_stride, int pair_off, int K,
                                         const float* sx0, int ldx, Pre pre, Epi epi, int units_per_group = 0x7fffffff,
                                         int sx_group_stride = 0, int row_group_stride = 0) {
    constexpr int VN = WVec<WT>::N;
    const int lane = threadIdx.x & 31;
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    const int GW = gridDim.x * T2S_WARPS;
    const int chunks = K / VN;
    const int half = t2s_half<WT>(K);
    if (t2s_dbg_skip_gemv) n_units = 0;
    for (int u = gw; u < n_units; u += GW) {
        const int grp = u / units_per_group;
        const int r0 = (u - grp * units_per_group) * unit_stride, r1 = r0 + pair_off;
        const int rofs = grp * row_group_stride;
        const float* sx = sx0 + grp * sx_group_stride;
        float2 pf = make_float2(0.f, 0.f);
        if (lane < NB) pf = pre(r0 + rofs, r1 + rofs, lane);
        const WT* w0p = W + static_cast<size_t>(r0) * ldw;
        const WT* w1p = W + static_cast<size_t>(r1) * ldw;
        float acc0[NB], acc1[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc0[b] = acc1[b] = 0.f;
#pragma unroll 4
        for (int c = lane; c < chunks; c += 32) {
            float w0[VN], w1[VN];
            WVec<WT>::load(w0p + c * VN, w0);
            if (R == 2) WVec<WT>::load(w1p + c * VN, w1);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
#pragma unroll
                for (int i = 0; i < VN; i += 4) {
                    const float4 xv = *reinterpret_cast<const float4*>(sx + b * ldx + c * 4 + (i ? half : 0));
                    acc0[b] = fmaf(w0[i], xv.x, acc0[b]);
                    acc0[b] = fmaf(w0[i + 1], xv.y, acc0[b]);
                    acc0[b] = fmaf(w0[i + 2], xv.z, acc0[b]);
                    acc0[b] = fmaf(w0[i + 3], xv.w, acc0[b]);
                    if (R == 2) {
                        acc1[b] = fmaf(w1[i], xv.x, acc1[b]);
                        acc1[b] = fmaf(w1[i + 1], xv.y, acc1[b]);
                        acc1[b] = fmaf(w1[i + 2], xv.z, acc1[b]);
                        acc1[b] = fmaf(w1[i + 3], xv.w, acc1[b]);
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc0[b] += __shfl_xor_sync(0xffffffffu, acc0[b], o);
                if (R == 2) acc1[b] += __shfl_xor_sync(0xffffffffu, acc1[b], o);
            }
        }
        if (lane < NB) epi(r0 + rofs, r1 + rofs, lane, t2s_select<NB>(acc0, lane), t2s_select<NB>(acc1, lane), pf);
    }
}

// A plain projection (no row pairing needed): two adjacent rows per unit when NB >= 4 (halves the shared-memory reads per
// FMA, which is what bounds the wide-batch stages), one row per unit otherwise (more warps, shorter chains).
// epi1(r, b, v, prefetched) / pre1(r, b) see single rows.
// n_groups > 1: the same n_rows weight rows applied to n_groups activation slices sx + g * sx_group_stride; rows are
// reported as g * n_rows + r.
template <class WT, int NB, class Pre1, class Epi1>
__device__ __forceinline__ void t2s_gemv_rows(const WT* __restrict__ W, int ldw, int n_rows, int K, const float* sx, int ldx,
                                              Pre1 pre1, Epi1 epi1, int n_groups = 1, int sx_group_stride = 0) {
    if (NB >= 4) {
        t2s_gemv<WT, NB, 2>(W, ldw, n_groups * (n_rows / 2), 2, 1, K, sx, ldx,
                            [=](int r0, int r1, int b) { return make_float2(pre1(r0, b), pre1(r1, b)); },
                            [=](int r0, int r1, int b, float v0, float v1, float2 pf) {
                                epi1(r0, b, v0, pf.x);
                                epi1(r1, b, v1, pf.y);
                            },
                            n_rows / 2, sx_group_stride, n_rows);
    } else {
        t2s_gemv<WT, NB, 1>(W, ldw, n_groups * n_rows, 1, 0, K, sx, ldx,
                            [=](int r0, int, int b) { return make_float2(pre1(r0, b), 0.f); },
                            [=](int r0, int, int b, float v0, float, float2 pf) { epi1(r0, b, v0, pf.x); }, n_rows,
                            sx_group_stride, n_rows);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core path for wide batches (bf16 matrices, NB >= 4).  With 8 rows of activations the CUDA-core product above is
// bound by shared-memory reads and FMA issue (16 LDS.128 + 128 FMA per 16-byte weight load); a warp-level
// mma.sync m16n8k16 takes the batch as its n = 8 dimension instead: A = 16 weight rows x 16 k (bf16, straight from global
// memory), B = 16 k x 8 batch rows (activations rounded to bf16 in registers), D = 16 x 8 fp32.  (tcgen05 wants M >= 64 rows of
// the *moving* operand per CTA and a TMEM round trip per stage; for a 16 x 8 x K product that is re-launched 26 times per
// decoding step behind a grid barrier, the register-resident warp MMA is the right tool.)
//  * k is permuted inside every 64-k block so that a lane's A fragments for four consecutive MMAs are two contiguous
//    16-byte loads per row (lane c of a quad owns k [8c, 8c+8) and [32+8c, 32+8c+8)); the lane fetches the matching
//    activations of its batch row (4 x 16 bytes of fp32, t2s_mma_kofs) straight from global memory next to its weight loads;
//  * a unit = 8 primary rows + 8 secondary rows (A rows 0-7 / 8-15), so a lane ends up with a (primary, secondary) pair for
//    two batch rows -- exactly what the rotary / GEGLU / two-plain-rows epilogues of the CUDA-core path take;
//  * the 16 warps of a CTA split K of ONE unit (one 64-k block each for K = 1024) and reduce their D fragments through
//    shared memory: all loads of a unit are in flight at once, and a stage of few units still finishes in one round trip.
__device__ __forceinline__ void t2s_mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// physical k (offset inside a 64-k block) of the 4 consecutive activations lane c of a quad feeds to MMA s
__device__ __forceinline__ int t2s_mma_kofs(int c, int s) { return (s < 2 ? 0 : 32) + 8 * c + 4 * (s & 1); }

// xg: fp32 activations [NB][K] in GLOBAL memory (written by other CTAs in the previous stage, read with ld.cg); gamma:
// RMSNorm gain [K] or nullptr.  A lane fetches exactly the 4 x 4 activations of its batch row that its MMAs need, next to
// its weight loads (no staging pass through shared memory, no extra block sync), multiplies by gamma and rounds to bf16.
// With gamma the row factor sqrt(K) / ||x|| of RMSNorm (text2semantic.py:143-151) is applied to the sums: the 16 warps of the
// CTA read every element of the rows exactly once per round, so sum(x^2) rides along with the fragment exchange.
template <int NB, int U, class Pre, class Epi>
__device__ __forceinline__ void t2s_gemv_mma_u(const __nv_bfloat16* __restrict__ W, int ldw, int n_units, int unit_rows,
                                               int row_stride, int pair_off, int primary_limit, int K, const float* xg,
                                               const float* __restrict__ gamma, float4* sfrag, float* sssq, Pre pre, Epi epi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int nblk = K / 64;
    const int nw = nblk < T2S_WARPS ? nblk : T2S_WARPS;
    // a CTA takes U units per round: U x 4 weight loads per lane in flight, and warps 0..U-1 reduce one unit each
    for (int u0 = blockIdx.x * U; u0 < n_units; u0 += gridDim.x * U) {
        float2 pf0 = make_float2(0.f, 0.f), pf1 = pf0;
        int my_r0 = 0;
        bool my_valid = false;
        if (warp < U && u0 + warp < n_units) {
            my_r0 = (u0 + warp) * unit_rows + g * row_stride;
            my_valid = my_r0 < primary_limit;
            if (my_valid) {
                if (2 * c < NB) pf0 = pre(my_r0, my_r0 + pair_off, 2 * c);
                if (2 * c + 1 < NB) pf1 = pre(my_r0, my_r0 + pair_off, 2 * c + 1);
            }
        }
        float d[U][4];
#pragma unroll
        for (int i = 0; i < U; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
        float ss = 0.f;
        for (int blk = warp; blk < nblk; blk += T2S_WARPS) {
            uint4 p0[U], p1[U], s0[U], s1[U];
#pragma unroll
            for (int i = 0; i < U; ++i) {
                int r0 = (u0 + i) * unit_rows + g * row_stride;
                if (u0 + i >= n_units || r0 >= primary_limit) r0 = 0;
                const __nv_bfloat16* pa = W + static_cast<size_t>(r0) * ldw + blk * 64 + 8 * c;
                const __nv_bfloat16* pb = W + static_cast<size_t>(r0 + pair_off) * ldw + blk * 64 + 8 * c;
                p0[i] = t2s_ld_stream(pa), p1[i] = t2s_ld_stream(pa + 32);
                s0[i] = t2s_ld_stream(pb), s1[i] = t2s_ld_stream(pb + 32);
            }
            float4 xv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                xv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g < NB) xv[q] = __ldcg(reinterpret_cast<const float4*>(xg + static_cast<size_t>(g) * K + blk * 64 + t2s_mma_kofs(c, q)));
            }
            uint32_t xb[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 v = xv[q];
                ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
                if (gamma != nullptr) {
                    const float4 gm = *reinterpret_cast<const float4*>(gamma + blk * 64 + t2s_mma_kofs(c, q));
                    v.x *= gm.x, v.y *= gm.y, v.z *= gm.z, v.w *= gm.w;
                }
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                xb[q][0] = *reinterpret_cast<const uint32_t*>(&lo);
                xb[q][1] = *reinterpret_cast<const uint32_t*>(&hi);
            }
#pragma unroll
            for (int i = 0; i < U; ++i) {
                t2s_mma_bf16(d[i], p0[i].x, s0[i].x, p0[i].y, s0[i].y, xb[0][0], xb[0][1]);
                t2s_mma_bf16(d[i], p0[i].z, s0[i].z, p0[i].w, s0[i].w, xb[1][0], xb[1][1]);
                t2s_mma_bf16(d[i], p1[i].x, s1[i].x, p1[i].y, s1[i].y, xb[2][0], xb[2][1]);
                t2s_mma_bf16(d[i], p1[i].z, s1[i].z, p1[i].w, s1[i].w, xb[3][0], xb[3][1]);
            }
        }
#pragma unroll
        for (int i = 0; i < U; ++i) sfrag[(i * T2S_WARPS + warp) * 32 + lane] = make_float4(d[i][0], d[i][1], d[i][2], d[i][3]);
        if (gamma != nullptr) {       // sum(x^2) of batch row g over this warp's blocks: reduce the quad, one slot per (warp, row)
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            if (c == 0) sssq[warp * 8 + g] = ss;
        }
        __syncthreads();
        if (warp < U && u0 + warp < n_units) {
            const float4* sf = sfrag + warp * T2S_WARPS * 32;
            float4 acc = sf[lane];
            for (int w = 1; w < nw; ++w) {
                const float4 v = sf[w * 32 + lane];
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            }
            if (gamma != nullptr) {
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int w = 0; w < T2S_WARPS; ++w) {
                    t0 += sssq[w * 8 + 2 * c];
                    t1 += sssq[w * 8 + 2 * c + 1];
                }
                const float sc0 = sqrtf(static_cast<float>(K)) / fmaxf(sqrtf(t0), 1e-12f);
                const float sc1 = sqrtf(static_cast<float>(K)) / fmaxf(sqrtf(t1), 1e-12f);
                acc.x *= sc0, acc.z *= sc0, acc.y *= sc1, acc.w *= sc1;
            }
            if (my_valid) {
                if (2 * c < NB) epi(my_r0, my_r0 + pair_off, 2 * c, acc.x, acc.z, pf0);
                if (2 * c + 1 < NB) epi(my_r0, my_r0 + pair_off, 2 * c + 1, acc.y, acc.w, pf1);
            }
        }
        __syncthreads();
    }
}

template <int NB, class Pre, class Epi>
__device__ __forceinline__ void t2s_gemv_mma(const __nv_bfloat16* __restrict__ W, int ldw, int n_units, int unit_rows,
                                             int row_stride, int pair_off, int primary_limit, int K, const float* xg,
                                             const float* __restrict__ gamma, float4* sfrag, float* sssq, Pre pre, Epi epi) {
    if (t2s_dbg_skip_gemv) return;
    const int per_cta = (n_units + gridDim.x - 1) / gridDim.x;
    if (per_cta <= 1) t2s_gemv_mma_u<NB, 1>(W, ldw, n_units, unit_rows, row_stride, pair_off, primary_limit, K, xg, gamma, sfrag, sssq, pre, epi);
    else if (per_cta <= 2) t2s_gemv_mma_u<NB, 2>(W, ldw, n_units, unit_rows, row_stride, pair_off, primary_limit, K, xg, gamma, sfrag, sssq, pre, epi);
    else t2s_gemv_mma_u<NB, 3>(W, ldw, n_units, unit_rows, row_stride, pair_off, primary_limit, K, xg, gamma, sfrag, sssq, pre, epi);
}

// L2 prefetch of the rows of this CTA's units in the next tensor-core stage
__device__ __forceinline__ void t2s_prefetch_team(const void* W, int row_bytes, int n_units, int unit_rows, int row_stride,
                                                  int pair_off, int primary_limit) {
    if (t2s_dbg_no_prefetch) return;
    const char* base = static_cast<const char*>(W);
    const int lines = (row_bytes + 127) / 128;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        for (int i = threadIdx.x; i < 16 * lines; i += T2S_THREADS) {
            const int rr = i / lines, ln = i - rr * lines;
            const int r0 = u * unit_rows + (rr & 7) * row_stride;
            if (r0 >= primary_limit) continue;
            const char* row = base + static_cast<size_t>(r0 + (rr >> 3) * pair_off) * row_bytes + ln * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
        }
    }
}

// Ask L2 for the weight rows this warp will use in the NEXT stage before waiting at the grid barrier: the rows do not
// depend on the activations, so an HBM miss overlaps the barrier and the activation load.  (L1 would be the better target
// but gpu-scope fences invalidate it.)
__device__ __forceinline__ void t2s_prefetch_rows(const void* W, int row_bytes, int n_units, int unit_stride, int pair_off,
                                                  int R) {
    if (t2s_dbg_no_prefetch) return;
    const int lane = threadIdx.x & 31;
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, GW = gridDim.x * T2S_WARPS;
    const char* base = static_cast<const char*>(W);
    for (int u = gw; u < n_units; u += GW) {
        for (int r = 0; r < R; ++r) {
            const char* row = base + static_cast<size_t>(u * unit_stride + r * pair_off) * row_bytes;
            for (int off = lane * 128; off < row_bytes; off += 32 * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
        }
    }
}
template <class WT, int NB>
__device__ __forceinline__ void t2s_prefetch_plain(const void* W, int ldw, int n_rows) {
    if (NB >= 4) t2s_prefetch_rows(W, ldw * static_cast<int>(sizeof(WT)), n_rows / 2, 2, 1, 2);
    else t2s_prefetch_rows(W, ldw * static_cast<int>(sizeof(WT)), n_rows, 1, 0, 1);
}

// sx[b][idx(j)] = x[b][j] * gamma[j];  sscale[b] = sqrt(D) / max(||x[b]||, 1e-12).  RMSNorm (text2semantic.py:143-151) is
// x / ||x|| * sqrt(D) * gamma; the per-row factor commutes with the matrix product that follows, so it is applied to the
// sums in the epilogue (no second pass over the row).  x is written by other CTAs -> L2 loads (__ldcg).  D % 4 == 0.
template <int NB>
__device__ __forceinline__ void t2s_load_norm(const float* x, const float* __restrict__ gamma, int D, int half, float* sx,
                                              int ldx, float* sred, float* sscale) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float ss[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) ss[b] = 0.f;
    for (int j = tid * 4; j < D; j += T2S_THREADS * 4) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + j);
        float4 v[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = __ldcg(reinterpret_cast<const float4*>(x + b * D + j));
        const int dst = t2s_sx_index(j, half);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            *reinterpret_cast<float4*>(sx + b * ldx + dst) = make_float4(v[b].x * g.x, v[b].y * g.y, v[b].z * g.z, v[b].w * g.w);
            ss[b] = fmaf(v[b].x, v[b].x, fmaf(v[b].y, v[b].y, fmaf(v[b].z, v[b].z, fmaf(v[b].w, v[b].w, ss[b]))));
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss[b] += __shfl_xor_sync(0xffffffffu, ss[b], o);
        if (lane == 0) sred[b * T2S_WARPS + warp] = ss[b];
    }
    __syncthreads();
    if (tid < NB) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < T2S_WARPS; ++w) t += sred[tid * T2S_WARPS + w];
        sscale[tid] = sqrtf(static_cast<float>(D)) / fmaxf(sqrtf(t), 1e-12f);
    }
    __syncthreads();
}

// sx[b][idx(j)] = src[b][j], j < n  (an activation other CTAs produced; n % 4 == 0)
template <int NB>
__device__ __forceinline__ void t2s_load_plain(const float* src, int n, int half, float* sx, int ldx) {
    const int n4 = n / 4;
#pragma unroll 4
    for (int i = threadIdx.x; i < NB * n4; i += T2S_THREADS) {
        const int b = i / n4, j = (i - b * n4) * 4;
        *reinterpret_cast<float4*>(sx + b * ldx + t2s_sx_index(j, half)) = __ldcg(reinterpret_cast<const float4*>(src + b * n + j));
    }
    __syncthreads();
}

// Single-query attention stage (attend_t2s.py:127-171 with one query row).  A CTA owns a unit = (b, h, group of up to 16
// consecutive 32-key blocks); warp w takes block w: a lane owns one key for the score (its 256-byte K row: 16 independent
// 16-byte loads), the warp takes max / sum with shuffles, then a lane owns two output dims and the 32 V rows are read
// coalesced with the probabilities broadcast by shuffle -- all loads of a block are independent, so a block costs about one
// memory round trip.  The warps' (m, l, o[64]) are merged in shared memory.  The group size is chosen so that the units just
// fill the grid; when one group covers all keys (always for the cross attention, and for the first 512 positions of the
// self attention at B = 8) the result goes straight to attn[b][h*64 ...]; otherwise the CTA writes a partial and the last CTA
// to arrive for (b, h) merges the partials -- no extra grid barrier either way.  Masked keys score -FLT_MAX like the
// reference's masked_fill (attend_t2s.py:151-153).
template <bool KV16>
__device__ __forceinline__ void t2s_attention_stage(const T2SDecArgs& a, int NB, const void* Kc, const void* Vc, int n_alloc,
                                                    int nkeys, const uint8_t* mask, float* sq, float* spart, int* sflag, int trace_base) {
#define T2S_AMARK(id)                                                                                  \
    do {                                                                                              \
        if (trace_base >= 0 && blockIdx.x == 0 && threadIdx.x == 0) t2s_trace_buf[trace_base + (id)] = clock64(); \
    } while (0)
    T2S_AMARK(0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblocks = (nkeys + 31) / 32;
    const int BH = NB * a.H;
    // groups: one CTA per (b, h) as long as a warp gets at most T2S_BLOCKS_PER_WARP blocks (2048 keys): the result goes
    // straight to attn[] with an in-smem merge -- no partials, fence, atomic or last-arriver merge through memory; longer
    // sequences are split into several groups
    int ngroups = (nblocks + T2S_BLOCKS_PER_WARP * T2S_WARPS - 1) / (T2S_BLOCKS_PER_WARP * T2S_WARPS);
    if (t2s_dbg_attn_groups > 0) ngroups = min(t2s_dbg_attn_groups, nblocks);
    const int bpg = (nblocks + ngroups - 1) / ngroups;
    const int units = BH * ngroups;
    constexpr float NEG = -3.402823466e38f;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int g = u % ngroups, bh = u / ngroups, b = bh / a.H;
        if (tid < T2S_DH / 2) {
            const float2 qv = __ldcg(reinterpret_cast<const float2*>(a.q + static_cast<size_t>(bh) * T2S_DH) + tid);
            sq[2 * tid] = qv.x * 0.125f;            // dim_head ** -0.5
            sq[2 * tid + 1] = qv.y * 0.125f;
        }
        __syncthreads();
        T2S_AMARK(1);
        float m = NEG, l = 0.f;
        float2 o = make_float2(0.f, 0.f);
        for (int wb = warp; wb < bpg; wb += T2S_WARPS) {
            const int kb = (g * bpg + wb) * 32;
            if (kb >= nkeys) break;
            const int j = kb + lane;
            const int cnt = min(32, nkeys - kb);
            float sc = NEG;
            uint32_t vw[KV16 ? 32 : 1];
            if constexpr (KV16) {
                // a lane owns key kb + lane: one coalesced 4-byte word (two dims) per load, 32 loads
                const uint32_t* Kw = static_cast<const uint32_t*>(Kc) + (static_cast<size_t>(bh) * (n_alloc / 32) + kb / 32) * 1024 + lane;
                uint32_t kw[32];
#pragma unroll
                for (int d = 0; d < 32; ++d) kw[d] = __ldcg(Kw + d * 32);
                // the V rows do not depend on the scores: request them now, one memory round trip for both
                const uint32_t* vr = static_cast<const uint32_t*>(Vc) + (static_cast<size_t>(bh) * n_alloc + kb) * 32 + lane;
#pragma unroll
                for (int t = 0; t < 32; ++t) vw[t] = t < cnt ? __ldcg(vr + t * 32) : 0u;
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < 32; ++d) {
                    const float2 qq = *reinterpret_cast<const float2*>(sq + 2 * d);
                    acc = fmaf(qq.x, __uint_as_float(kw[d] << 16), acc);
                    acc = fmaf(qq.y, __uint_as_float(kw[d] & 0xffff0000u), acc);
                }
                if (j < nkeys) sc = (mask != nullptr && !mask[b * a.n_ctx + j]) ? NEG : acc;
            } else {
                const float* Kb = static_cast<const float*>(Kc) + static_cast<size_t>(bh) * n_alloc * T2S_DH;
                if (j < nkeys) {
                    const float4* kr = reinterpret_cast<const float4*>(Kb + static_cast<size_t>(j) * T2S_DH);
                    float4 kv[T2S_DH / 4];
#pragma unroll
                    for (int d = 0; d < T2S_DH / 4; ++d) kv[d] = __ldcg(kr + d);
                    float acc = 0.f;
#pragma unroll
                    for (int d = 0; d < T2S_DH / 4; ++d) {
                        const float4 qq = *reinterpret_cast<const float4*>(sq + 4 * d);
                        acc = fmaf(qq.x, kv[d].x, acc);
                        acc = fmaf(qq.y, kv[d].y, acc);
                        acc = fmaf(qq.z, kv[d].z, acc);
                        acc = fmaf(qq.w, kv[d].w, acc);
                    }
                    sc = (mask != nullptr && !mask[b * a.n_ctx + j]) ? NEG : acc;
                }
            }
            T2S_AMARK(2);
            float mb = sc;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, off));
            const float p = j < nkeys ? expf(sc - mb) : 0.f;
            float lb = p;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) lb += __shfl_xor_sync(0xffffffffu, lb, off);
            T2S_AMARK(3);
            float2 ob = make_float2(0.f, 0.f);
            if constexpr (KV16) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const float pt = __shfl_sync(0xffffffffu, p, t);
                    ob.x = fmaf(pt, __uint_as_float(vw[t] << 16), ob.x);
                    ob.y = fmaf(pt, __uint_as_float(vw[t] & 0xffff0000u), ob.y);
                }
            } else {
                const float* Vb = static_cast<const float*>(Vc) + static_cast<size_t>(bh) * n_alloc * T2S_DH;
                const float2* vr = reinterpret_cast<const float2*>(Vb + static_cast<size_t>(kb) * T2S_DH) + lane;
                if (cnt == 32) {
                    float2 vv[32];
#pragma unroll
                    for (int t = 0; t < 32; ++t) vv[t] = __ldcg(vr + t * (T2S_DH / 2));
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        const float pt = __shfl_sync(0xffffffffu, p, t);
                        ob.x = fmaf(pt, vv[t].x, ob.x);
                        ob.y = fmaf(pt, vv[t].y, ob.y);
                    }
                } else {
                    for (int t = 0; t < cnt; ++t) {
                        const float2 vv = __ldcg(vr + t * (T2S_DH / 2));
                        const float pt = __shfl_sync(0xffffffffu, p, t);
                        ob.x = fmaf(pt, vv.x, ob.x);
                        ob.y = fmaf(pt, vv.y, ob.y);
                    }
                }
            }
            // online merge of this warp's blocks
            const float mn = fmaxf(m, mb);
            const float ra = l > 0.f ? expf(m - mn) : 0.f, rb = expf(mb - mn);
            l = fmaf(l, ra, lb * rb);
            o.x = fmaf(o.x, ra, ob.x * rb);
            o.y = fmaf(o.y, ra, ob.y * rb);
            m = mn;
        }
        T2S_AMARK(4);
        float* sp = spart + warp * T2S_PART;
        if (lane == 0) sp[0] = m, sp[1] = l;
        *reinterpret_cast<float2*>(sp + 2 + 2 * lane) = o;
        __syncthreads();
        T2S_AMARK(5);
        // merge the warps' partials: thread d < 64 owns output dim d
        float Mg = NEG, Lg = 0.f, og = 0.f;
        const int nw = min(min(bpg, nblocks - g * bpg), T2S_WARPS);
        if (tid < T2S_DH) {
            for (int w = 0; w < nw; ++w) Mg = fmaxf(Mg, spart[w * T2S_PART]);
            for (int w = 0; w < nw; ++w) {
                const float lw = spart[w * T2S_PART + 1];
                const float wt = lw > 0.f ? expf(spart[w * T2S_PART] - Mg) : 0.f;
                Lg = fmaf(lw, wt, Lg);
                og = fmaf(spart[w * T2S_PART + 2 + tid], wt, og);
            }
        }
        if (ngroups == 1) {
            if (tid < T2S_DH) a.attn[static_cast<size_t>(bh) * T2S_DH + tid] = og / Lg;
            __syncthreads();
            T2S_AMARK(6);
            continue;
        }
        float* part = a.part + (static_cast<size_t>(bh) * T2S_MAXS + g) * T2S_PART;
        if (tid < T2S_DH) {
            part[2 + tid] = og;
            if (tid == 0) part[0] = Mg, part[1] = Lg;
            __threadfence();
        }
        __syncthreads();
        T2S_AMARK(7);
        if (tid == 0) sflag[0] = atomicAdd(a.part_cnt + bh, 1u) == static_cast<unsigned>(ngroups - 1);
        __syncthreads();
        T2S_AMARK(8);
        if (sflag[0] && tid < T2S_DH) {
            __threadfence();
            const float* pb = a.part + static_cast<size_t>(bh) * T2S_MAXS * T2S_PART;
            // all loads first (independent), then the max / weighted sum
            float M2 = NEG, L2 = 0.f, o2 = 0.f;
            for (int t0 = 0; t0 < ngroups; t0 += 8) {
                float mt[8], lt[8], pv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool ok = t0 + i < ngroups;
                    const float* pp = pb + (ok ? t0 + i : t0) * T2S_PART;
                    mt[i] = __ldcg(pp);
                    lt[i] = ok ? __ldcg(pp + 1) : 0.f;
                    pv[i] = __ldcg(pp + 2 + tid);
                }
                float Mn = M2;
#pragma unroll
                for (int i = 0; i < 8; ++i) Mn = lt[i] > 0.f ? fmaxf(Mn, mt[i]) : Mn;
                const float r = L2 > 0.f ? expf(M2 - Mn) : 0.f;
                L2 *= r;
                o2 *= r;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float wt = lt[i] > 0.f ? expf(mt[i] - Mn) : 0.f;
                    L2 = fmaf(lt[i], wt, L2);
                    o2 = fmaf(pv[i], wt, o2);
                }
                M2 = Mn;
            }
            a.attn[static_cast<size_t>(bh) * T2S_DH + tid] = o2 / L2;
            if (tid == 0) a.part_cnt[bh] = 0u;       // next use is behind at least one grid barrier
        }
        __syncthreads();
        T2S_AMARK(9);
    }
#undef T2S_AMARK
}

// Monotone map float -> uint32 (larger float <=> larger key), for the radix select below
__device__ __forceinline__ unsigned t2s_sortable(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// top-k filter + Gumbel arg-max for one (stream, batch row) (text2semantic.py:104-132, :793-800); also feeds the next
// position: x[b][s*demb ...] = emb[token]  (:746-751).  The k-th largest logit is found by a 4-pass radix select over the
// order-preserving integer image of the logits (shared-memory histograms); logits >= it survive the filter (torch.topk keeps
// exactly k: identical unless two logits tie bit-for-bit at the threshold).
__device__ __forceinline__ void t2s_sample_unit(const T2SDecArgs& a, int s, int b, int step, unsigned* hist, float* sval,
                                                int* sidx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n_logits;                         // <= T2S_THREADS (checked on the host)
    const size_t off = (static_cast<size_t>(s) * a.B + b) * n;
    float l = -INFINITY, g = 0.f;
    long long fed = -1;
    if (tid < n) {
        l = __ldcg(a.logits + off + tid);
        const float uu = a.u[static_cast<size_t>(step) * a.n_out * a.B * n + off + tid];
        g = -logf(fmaxf(-logf(fmaxf(uu, 1e-20f)), 1e-20f));
        if (a.logits_out != nullptr) a.logits_out[static_cast<size_t>(step) * a.n_out * a.B * n + off + tid] = l;
    }
    if (tid == 0 && a.forced != nullptr) fed = a.forced[(static_cast<size_t>(b) * a.n_out + s) * a.max_len + step];
    const unsigned key = tid < n ? t2s_sortable(l) : 0u;
    unsigned prefix = 0u, pmask = 0u;
    int want = a.topk;                                // rank (1-based, from the top) still to locate inside the prefix class
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        if (tid < n && (key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        __syncthreads();
        if (warp == 0) {                              // find the bin (from the top) where the cumulative count reaches `want`
            unsigned c[8], tot = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                c[i] = hist[255 - (lane * 8 + i)];
                tot += c[i];
            }
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            unsigned run = incl - tot;                // elements in higher bins
            int found = -1, rem = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (found < 0 && run + c[i] >= static_cast<unsigned>(want)) {
                    found = 255 - (lane * 8 + i);
                    rem = want - static_cast<int>(run);
                }
                run += c[i];
            }
            const unsigned ball = __ballot_sync(0xffffffffu, found >= 0);
            const int src = __ffs(ball) - 1;          // lowest lane = highest bins
            found = __shfl_sync(0xffffffffu, found, src);
            rem = __shfl_sync(0xffffffffu, rem, src);
            if (lane == 0) sidx[0] = found, sidx[1] = rem;
        }
        __syncthreads();
        prefix |= static_cast<unsigned>(sidx[0]) << shift;
        pmask |= 255u << shift;
        want = sidx[1];
        __syncthreads();
    }
    // prefix == key of the k-th largest logit
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    if (tid < n && key >= prefix) best = l / fmaxf(a.temperature, 1e-10f) + g, best_i = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) best = ov, best_i = oi;
    }
    if (lane == 0) sval[warp] = best, sidx[2 + warp] = best_i;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < T2S_WARPS; ++w)
            if (sval[w] > best || (sval[w] == best && sidx[2 + w] < best_i)) best = sval[w], best_i = sidx[2 + w];
        if (best_i >= n) best_i = 0;               // only reachable with non-finite logits
        a.tokens[(static_cast<size_t>(b) * a.n_out + s) * a.max_len + step] = best_i;
        if (best_i == a.eos_id && !a.ignore_eos) a.eos_flags[s * a.B + b] = 1;
        if (a.forced == nullptr) fed = best_i;
        if (fed < 0 || fed >= n) fed = a.eos_id;   // pad (-1) after EOS: never reached for B = 1 (see t2s.py)
        sidx[0] = static_cast<int>(fed);
    }
    __syncthreads();
    const float* e = a.emb + static_cast<size_t>(sidx[0]) * a.demb;
    for (int j = tid; j < a.demb; j += T2S_THREADS) a.x[static_cast<size_t>(b) * a.Dt + s * a.demb + j] = e[j];
    __syncthreads();
}

template <class WT, int NB>
__global__ void __launch_bounds__(T2S_THREADS, 1) t2s_decode_kernel(const __grid_constant__ T2SDecArgs a) {
    constexpr bool kMMA = sizeof(WT) == 2;                 // bf16 matrices -> tensor-core path (mma.sync, batch = n dimension)
    constexpr bool kKV16 = sizeof(WT) == 2;                // bf16 matrices -> bf16 KV cache / context (layouts above)
    extern __shared__ __align__(16) float smem[];
    const int kmax = a.ffi_pad > a.Dt ? a.ffi_pad : a.Dt;
    const int ldx = kmax;                              // fp32 activation rows (CUDA-core path, and the logit stage)
    float* sx = smem;                                  // [NB][ldx] fp32 (CUDA-core path and the logit stage)
    float* spart = sx + NB * ldx;                      // 3 x [16][32] float4: attention partials, sampler histogram, MMA fragments
    float* sq = spart + 3 * T2S_WARPS * 32 * 4;        // [64]
    float* sred = sq + T2S_DH;                         // [NB * 16]
    float* sscale = sred + NB * T2S_WARPS;             // [8]
    int* sidx = reinterpret_cast<int*>(sscale + 8);    // [2 + 16 + 2]
    float* sssq = reinterpret_cast<float*>(sidx + 20); // [16 warps][8 rows] sum(x^2) partials (tensor-core path)
    float4* sfrag = reinterpret_cast<float4*>(spart);
    unsigned epoch = 0;
    const int tid = threadIdx.x;
    const int Dt = a.Dt, inner = a.inner;
    float* const x = a.x;
    const auto no_pre = [](int, int, int) { return make_float2(0.f, 0.f); };
    const auto no_pre1 = [](int, int) { return 0.f; };
    const auto resid = [=](int r, int b) { return __ldcg(x + b * Dt + r); };

    // activation staging for a projection with K inputs: RMSNorm'ed (gamma != nullptr) or plain
    const float* cur_src = nullptr;                    // tensor-core path: the projection reads its activations itself
    const float* cur_gamma = nullptr;
    auto stage_in = [&](const float* src, const float* gamma, int K) {
        if constexpr (kMMA) {
            cur_src = src;
            cur_gamma = gamma;
            (void)K;
        } else {
            if (gamma != nullptr) t2s_load_norm<NB>(src, gamma, K, t2s_half<WT>(K), sx, ldx, sred, sscale);
            else t2s_load_plain<NB>(src, K, t2s_half<WT>(K), sx, ldx);
        }
    };
    // a plain projection: n_rows rows of W [n_rows, K]; epi1(r, b, v, prefetched), pre1(r, b)
    auto proj_rows = [&](const void* W, int n_rows, int K, auto pre1, auto epi1) {
        if constexpr (kMMA) {
            t2s_gemv_mma<NB>(static_cast<const __nv_bfloat16*>(W), K, n_rows / 16, 16, 1, 8, n_rows, K, cur_src, cur_gamma, sfrag, sssq,
                             [=](int r0, int r1, int b) { return make_float2(pre1(r0, b), pre1(r1, b)); },
                             [=](int r0, int r1, int b, float v0, float v1, float2 pf) {
                                 epi1(r0, b, v0, pf.x);
                                 epi1(r1, b, v1, pf.y);
                             });
        } else {
            t2s_gemv_rows<WT, NB>(static_cast<const WT*>(W), K, n_rows, K, sx, ldx, pre1, epi1);
        }
    };
    auto prefetch_rows_plain = [&](const void* W, int K, int n_rows) {
        if constexpr (kMMA) t2s_prefetch_team(W, K * 2, n_rows / 16, 16, 1, 8, n_rows);
        else t2s_prefetch_plain<WT, NB>(W, K, n_rows);
    };
    auto prefetch_qkv = [&](const void* W) {
        if constexpr (kMMA) t2s_prefetch_team(W, Dt * 2, 3 * inner / 16, 16, 2, 1, 3 * inner);
        else t2s_prefetch_rows(W, Dt * static_cast<int>(sizeof(WT)), 3 * inner / 2, 2, 1, 2);
    };

    if constexpr (kMMA) {      // the tensor-core projections apply the RMSNorm row factor themselves: the epilogues' sscale[b] is 1
        if (tid < 8) sscale[tid] = 1.0f;
        __syncthreads();
    }
    // position 0 input: the start token (text2semantic.py:746-751)
    for (int i = blockIdx.x * T2S_THREADS + tid; i < NB * Dt; i += gridDim.x * T2S_THREADS) a.x[i] = a.start[i % Dt];
    if (t2s_grid_barrier(a, epoch)) return;

    for (int step = 0; step < a.max_len; ++step) {
        for (int L = 0; L < a.depth; ++L) {
            const T2SLayerW& w = a.L[L];
            const bool work = (a.dbg_mode & 1) == 0;
            const int mk = 2 + L * 24;
            T2S_MARK(mk + 0);
            // ---- S1: self-attention q | k | v of the new position; rotary at position `step`; k, v appended to the cache
            if (work) {
                stage_in(a.x, w.sa_gamma, Dt);
                T2S_MARK(mk + 1);
                const float2* rope = a.rope + static_cast<size_t>(step) * (T2S_DH / 2);
                const int H = a.H;
                float* qb = a.q;
                void* kc = w.kcache;
                void* vc = w.vcache;
                const int max_len_r = a.max_len_r;
                const auto epi = [=](int r0, int, int b, float v0, float v1, float2) {
                    const int sect = r0 / inner, c = r0 % inner, h = c / T2S_DH, d = c % T2S_DH;
                    v0 *= sscale[b];
                    v1 *= sscale[b];
                    if (sect < 2) {
                        const float2 cs = rope[d >> 1];
                        const float t0 = v0 * cs.x - v1 * cs.y;
                        v1 = v1 * cs.x + v0 * cs.y;
                        v0 = t0;
                    }
                    const size_t bh = static_cast<size_t>(b) * H + h;
                    if (sect == 0) {
                        *reinterpret_cast<float2*>(qb + b * inner + c) = make_float2(v0, v1);
                    } else if constexpr (kKV16) {
                        const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
                        const uint32_t word = *reinterpret_cast<const uint32_t*>(&pk);
                        if (sect == 1)      // blocked-transposed keys: [block of 32 positions][dim pair][position in block]
                            static_cast<uint32_t*>(kc)[(bh * (max_len_r / 32) + step / 32) * 1024 + (d >> 1) * 32 + (step & 31)] = word;
                        else
                            static_cast<uint32_t*>(vc)[(bh * max_len_r + step) * 32 + (d >> 1)] = word;
                    } else {
                        float* dst = static_cast<float*>(sect == 1 ? kc : vc) + (bh * max_len_r + step) * T2S_DH + d;
                        *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                    }
                };
                if constexpr (kMMA)
                    t2s_gemv_mma<NB>(static_cast<const __nv_bfloat16*>(w.sa_qkv), Dt, 3 * inner / 16, 16, 2, 1, 3 * inner, Dt, cur_src,
                                     cur_gamma, sfrag, sssq, no_pre, epi);
                else
                    t2s_gemv<WT, NB, 2>(static_cast<const WT*>(w.sa_qkv), Dt, 3 * inner / 2, 2, 1, Dt, sx, ldx, no_pre, epi);
                prefetch_rows_plain(w.sa_out, inner, Dt);
            }
            T2S_MARK(mk + 2);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 3);
            // ---- S2: causal self-attention of the one query over the step + 1 cached keys
            if (work && !(a.dbg_mode & 2))
                t2s_attention_stage<kKV16>(a, NB, w.kcache, w.vcache, a.max_len_r, step + 1, nullptr, sq, spart, sidx,
                                    (L == 0 && step == t2s_trace_step) ? 106 : -1);
            T2S_MARK(mk + 4);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 5);
            // ---- S3: to_out + residual
            if (work) {
                stage_in(a.attn, nullptr, inner);
                T2S_MARK(mk + 6);
                proj_rows(w.sa_out, Dt, inner, resid, [=](int r, int b, float v, float res) { x[b * Dt + r] = v + res; });
                prefetch_rows_plain(w.ca_q, Dt, inner);
            }
            T2S_MARK(mk + 7);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 8);
            // ---- S4: cross-attention query
            if (work) {
                stage_in(a.x, w.ca_gamma, Dt);
                float* qb = a.q;
                proj_rows(w.ca_q, inner, Dt, no_pre1, [=](int r, int b, float v, float) { qb[b * inner + r] = v * sscale[b]; });
                prefetch_rows_plain(w.ca_out, inner, Dt);
            }
            T2S_MARK(mk + 9);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 10);
            // ---- S5: cross attention over [null kv | encoded text] with the source padding mask
            if (work && !(a.dbg_mode & 2))
                t2s_attention_stage<kKV16>(a, NB, w.ctx_k, w.ctx_v, a.n_ctx_r, a.n_ctx, a.ctx_mask, sq, spart, sidx,
                                    (L == 0 && step == t2s_trace_step) ? 116 : -1);
            T2S_MARK(mk + 11);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 12);
            // ---- S6: to_out + residual
            if (work) {
                stage_in(a.attn, nullptr, inner);
                proj_rows(w.ca_out, Dt, inner, resid, [=](int r, int b, float v, float res) { x[b * Dt + r] = v + res; });
                if constexpr (kMMA) t2s_prefetch_team(w.ff1, Dt * 2, (a.ffi + 7) / 8, 8, 1, a.ffi, a.ffi);
                else t2s_prefetch_rows(w.ff1, Dt * static_cast<int>(sizeof(WT)), a.ffi, 1, a.ffi, 2);
            }
            T2S_MARK(mk + 13);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 14);
            // ---- S7: FF in-projection + GEGLU: row i (x part) is paired with row ffi + i (gate)
            if (work) {
                stage_in(a.x, w.ff_gamma, Dt);
                T2S_MARK(mk + 15);
                float* hb = a.hbuf;
                const float* b1 = w.ff1_b;
                const int ffi_pad = a.ffi_pad;
                const auto pre = [=](int r0, int r1, int) { return make_float2(b1[r0], b1[r1]); };
                const auto epi = [=](int r0, int, int b, float v0, float v1, float2 pf) {
                    hb[b * ffi_pad + r0] = gelu_erf(fmaf(v1, sscale[b], pf.y)) * fmaf(v0, sscale[b], pf.x);
                };
                if constexpr (kMMA)
                    t2s_gemv_mma<NB>(static_cast<const __nv_bfloat16*>(w.ff1), Dt, (a.ffi + 7) / 8, 8, 1, a.ffi, a.ffi, Dt, cur_src,
                                     cur_gamma, sfrag, sssq, pre, epi);
                else
                    t2s_gemv<WT, NB, 2>(static_cast<const WT*>(w.ff1), Dt, a.ffi, 1, a.ffi, Dt, sx, ldx, pre, epi);
                prefetch_rows_plain(w.ff2, ffi_pad, Dt);
            }
            T2S_MARK(mk + 16);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 17);
            // ---- S8: FF out-projection + bias + residual
            if (work) {
                stage_in(a.hbuf, nullptr, a.ffi_pad);
                T2S_MARK(mk + 18);
                const float* b2 = w.ff2_b;
                proj_rows(w.ff2, Dt, a.ffi_pad, [=](int r, int b) { return __ldcg(x + b * Dt + r) + b2[r]; },
                          [=](int r, int b, float v, float res) { x[b * Dt + r] = v + res; });
                if (L + 1 < a.depth) prefetch_qkv(a.L[L + 1].sa_qkv);
                else t2s_prefetch_plain<float, NB>(a.emb, a.demb, a.n_logits);
            }
            T2S_MARK(mk + 19);
            if (t2s_grid_barrier(a, epoch)) return;
            T2S_MARK(mk + 20);
        }
        T2S_MARK(100);
        // ---- S9: final norm + tied logit projection per output stream (text2semantic.py:762-776), fp32 table, fp32 activations
        if ((a.dbg_mode & 1) == 0) {
            t2s_load_norm<NB>(a.x, a.final_gamma, Dt, 0, sx, ldx, sred, sscale);
            {       // both output streams in one pass: group s reads the s-th half of the normalised row
                float* lg = a.logits;
                const int n = a.n_logits;
                t2s_gemv_rows<float, NB>(a.emb, a.demb, a.n_logits, a.demb, sx, ldx, no_pre1,
                                         [=](int vr, int b, float v, float) {
                                             const int s = vr / n, r = vr - s * n;
                                             lg[(static_cast<size_t>(s) * NB + b) * n + r] = v * sscale[b];
                                         },
                                         a.n_out, a.demb);
            }
            prefetch_qkv(a.L[0].sa_qkv);
            if constexpr (kMMA) {      // the fp32 logit stage wrote its row factors: back to 1 for the tensor-core stages
                __syncthreads();
                if (tid < 8) sscale[tid] = 1.0f;
                __syncthreads();
            }
        }
        T2S_MARK(101);
        if (t2s_grid_barrier(a, epoch)) return;
        T2S_MARK(102);
        // ---- S10: sampling, one CTA per (stream, row); writes the next position's input embedding
        if (!(a.dbg_mode & 8))
            for (int u = blockIdx.x; u < a.n_out * NB; u += gridDim.x)
                t2s_sample_unit(a, u / NB, u % NB, step, reinterpret_cast<unsigned*>(spart), sred, sidx);
        T2S_MARK(103);
        if (t2s_grid_barrier(a, epoch)) return;
        T2S_MARK(104);
        // ---- EOS logic (text2semantic.py:804-826): stop when every row of stream 1 -- or, with two outputs, every row
        // of either stream -- has produced an EOS
        bool stop = false;
        for (int s = 0; s < a.n_out; ++s) {
            bool all = true;
            for (int b = 0; b < NB; ++b) all = all && (__ldcg(a.eos_flags + s * NB + b) != 0);
            stop = stop || all;
        }
        if (blockIdx.x == 0 && tid == 0) {
            a.result[0] = step + 1;
            a.result[1] = stop ? 1 : 0;
        }
        if (stop) break;
    }
}

struct T2SEncLayerW {
    Tensor attn_gamma, q_w, kv_w, out_w, ff_gamma, ff1_w, ff1_b, ff2_w, ff2_b;
};
struct T2SDecLayerW {
    Tensor sa_gamma, sa_qkv, sa_out, ca_gamma, ca_q, ca_kv, ca_null, ca_out, ff_gamma, ff1_w, ff1_b, ff2_w, ff2_b;
};

}  // namespace covo

struct covo_t2s {
    covo_t2s_cfg cfg;
    covo::DeviceInfo di;
    covo::Weights w;
    covo::Tensor enc_emb, enc_final, inv_freq, dec_emb, dec_start, dec_final;
    std::vector<covo::T2SEncLayerW> enc;
    std::vector<covo::T2SDecLayerW> dec;
    int inner = 0, fi_enc = 0, ffi = 0, ffi_pad = 0, n_out = 1, demb = 0, n_logits = 0;
    uint32_t wdt = covo::DT_BF16;
    int grid = 0;
};

namespace covo {

inline int t2s_bind_weights(covo_t2s* h) {
    const covo_t2s_cfg& c = h->cfg;
    const Weights& w = h->w;
    h->wdt = c.weight_format == COVO_T2S_W_F32 ? DT_F32 : DT_BF16;
    h->inner = c.heads * c.dim_head;
    h->fi_enc = static_cast<int>(c.dim * c.ff_mult * 2 / 3);
    h->ffi = static_cast<int>(c.target_transformer_dim * c.ff_mult * 2 / 3);
    h->ffi_pad = round_up(h->ffi, 64);      // K of the FF out-projection: whole 64-k blocks (tensor-core path)
    h->n_out = c.two_output ? 2 : 1;
    h->demb = c.target_transformer_dim / h->n_out;
    h->n_logits = c.num_semantic_token_ids + 1;
    COVO_TRY(w.get("enc.emb", DT_F32, &h->enc_emb));
    COVO_TRY(w.get("enc.final.gamma", DT_F32, &h->enc_final));
    COVO_TRY(w.get("rope.inv_freq", DT_F32, &h->inv_freq));
    COVO_TRY(w.get("dec.emb", DT_F32, &h->dec_emb));
    COVO_TRY(w.get("dec.start", DT_F32, &h->dec_start));
    COVO_TRY(w.get("dec.final.gamma", DT_F32, &h->dec_final));
    if (static_cast<int>(h->dec_emb.shape[0]) != h->n_logits || static_cast<int>(h->dec_emb.shape[1]) != h->demb)
        return fail(COVO_ERR_WEIGHTS, "dec.emb is [%llu, %llu], expected [%d, %d]", (unsigned long long)h->dec_emb.shape[0],
                    (unsigned long long)h->dec_emb.shape[1], h->n_logits, h->demb);
    if (static_cast<int>(h->enc_emb.shape[0]) != c.num_text_token_ids + 1 || static_cast<int>(h->enc_emb.shape[1]) != c.dim)
        return fail(COVO_ERR_WEIGHTS, "enc.emb has unexpected shape");
    h->enc.resize(c.source_depth);
    for (int L = 0; L < c.source_depth; ++L) {
        T2SEncLayerW& e = h->enc[L];
        const std::string p = "enc.L" + std::to_string(L) + ".";
        COVO_TRY(w.get(p + "attn.gamma", DT_F32, &e.attn_gamma));
        COVO_TRY(w.get(p + "q.w", DT_F32, &e.q_w));
        COVO_TRY(w.get(p + "kv.w", DT_F32, &e.kv_w));
        COVO_TRY(w.get(p + "out.w", DT_F32, &e.out_w));
        COVO_TRY(w.get(p + "ff.gamma", DT_F32, &e.ff_gamma));
        COVO_TRY(w.get(p + "ff1.w", DT_F32, &e.ff1_w));
        COVO_TRY(w.get(p + "ff1.b", DT_F32, &e.ff1_b));
        COVO_TRY(w.get(p + "ff2.w", DT_F32, &e.ff2_w));
        COVO_TRY(w.get(p + "ff2.b", DT_F32, &e.ff2_b));
    }
    h->dec.resize(c.target_depth);
    for (int L = 0; L < c.target_depth; ++L) {
        T2SDecLayerW& d = h->dec[L];
        const std::string p = "dec.L" + std::to_string(L) + ".";
        COVO_TRY(w.get(p + "sa.gamma", DT_F32, &d.sa_gamma));
        COVO_TRY(w.get(p + "sa.qkv.w", h->wdt, &d.sa_qkv));
        COVO_TRY(w.get(p + "sa.out.w", h->wdt, &d.sa_out));
        COVO_TRY(w.get(p + "ca.gamma", DT_F32, &d.ca_gamma));
        COVO_TRY(w.get(p + "ca.q.w", h->wdt, &d.ca_q));
        COVO_TRY(w.get(p + "ca.kv.w", DT_F32, &d.ca_kv));
        COVO_TRY(w.get(p + "ca.null_kv", DT_F32, &d.ca_null));
        COVO_TRY(w.get(p + "ca.out.w", h->wdt, &d.ca_out));
        COVO_TRY(w.get(p + "ff.gamma", DT_F32, &d.ff_gamma));
        COVO_TRY(w.get(p + "ff1.w", h->wdt, &d.ff1_w));
        COVO_TRY(w.get(p + "ff1.b", DT_F32, &d.ff1_b));
        COVO_TRY(w.get(p + "ff2.w", h->wdt, &d.ff2_w));
        COVO_TRY(w.get(p + "ff2.b", DT_F32, &d.ff2_b));
        if (static_cast<int>(d.ff2_w.shape[1]) != h->ffi_pad || static_cast<int>(d.ff1_w.shape[0]) != 2 * h->ffi)
            return fail(COVO_ERR_WEIGHTS, "dec.L%d FF weights have unexpected shape", L);
    }
    return COVO_OK;
}

inline int t2s_pad_batch(int B) { return B <= 1 ? 1 : (B <= 2 ? 2 : (B <= 4 ? 4 : 8)); }

inline size_t t2s_decode_smem(const covo_t2s* h, int NB) {
    const int kmax = h->ffi_pad > h->cfg.target_transformer_dim ? h->ffi_pad : h->cfg.target_transformer_dim;
    return sizeof(float) * (static_cast<size_t>(NB) * kmax + 3 * T2S_WARPS * 32 * 4 + T2S_DH + NB * T2S_WARPS + 8 + 20 + T2S_WARPS * 8);
}

// Workspace layout for (B rows padded to NB, S1 text positions incl. EOS, max_len decode positions)
struct T2SBuffers {
    // source side
    long long* ids;
    uint8_t *mask, *cmask;
    float *xe, *he, *qe, *kve, *ae, *f1e, *ge, *tmp;
    float2* rope;
    // decode
    float *ctx_k, *ctx_v, *kcache, *vcache, *x, *q, *attn, *part, *hbuf, *logits;
    int *result, *eos_flags;
    unsigned *barrier, *part_cnt;
    size_t zero_from = 0, zero_bytes = 0;     // region cleared before every call
};

inline size_t t2s_layout(const covo_t2s* h, int NB, int S1, int max_len, void* ws, T2SBuffers* b) {
    const covo_t2s_cfg& c = h->cfg;
    Arena a(ws, ~static_cast<size_t>(0));
    const size_t M = static_cast<size_t>(NB) * S1;
    const int n_ctx = S1 + 1, H = c.heads;
    const int rope_n = max_len > S1 ? max_len : S1;
    T2SBuffers t;
    t.ids = a.take<long long>(M);
    t.mask = a.take<uint8_t>(M);
    t.cmask = a.take<uint8_t>(static_cast<size_t>(NB) * n_ctx);
    t.xe = a.take<float>(M * c.dim);
    t.he = a.take<float>(M * c.dim);
    t.qe = a.take<float>(M * h->inner);
    t.kve = a.take<float>(M * 2 * h->inner);
    t.ae = a.take<float>(M * h->inner);
    t.f1e = a.take<float>(M * 2 * h->fi_enc);
    t.ge = a.take<float>(M * h->fi_enc);
    t.tmp = a.take<float>(M * c.dim);
    t.rope = a.take<float2>(static_cast<size_t>(rope_n) * (T2S_DH / 2));
    const size_t ctx_n = static_cast<size_t>(c.target_depth) * NB * H * round_up(n_ctx, 32) * T2S_DH;      // fp32-sized: covers both layouts
    const size_t cache_n = static_cast<size_t>(c.target_depth) * NB * H * round_up(max_len, 32) * T2S_DH;
    t.ctx_k = a.take<float>(ctx_n);
    t.ctx_v = a.take<float>(ctx_n);
    t.kcache = a.take<float>(cache_n);
    t.vcache = a.take<float>(cache_n);
    t.x = a.take<float>(static_cast<size_t>(NB) * c.target_transformer_dim);
    t.q = a.take<float>(static_cast<size_t>(NB) * h->inner);
    t.attn = a.take<float>(static_cast<size_t>(NB) * h->inner);
    t.part = a.take<float>(static_cast<size_t>(NB) * H * T2S_MAXS * T2S_PART);
    t.logits = a.take<float>(static_cast<size_t>(h->n_out) * NB * h->n_logits);
    a.off = align_up(a.off, 256);
    t.zero_from = a.off;
    t.hbuf = a.take<float>(static_cast<size_t>(NB) * h->ffi_pad);
    t.result = a.take<int>(4);
    t.eos_flags = a.take<int>(static_cast<size_t>(h->n_out) * NB);
    t.barrier = a.take<unsigned>(4);
    t.part_cnt = a.take<unsigned>(static_cast<size_t>(NB) * H);
    t.zero_bytes = a.off - t.zero_from;
    if (b) *b = t;
    return align_up(a.off, 256);
}

template <class WT, int NB>
inline int t2s_launch_decode(const covo_t2s* h, const T2SDecArgs& args, size_t smem, cudaStream_t st) {
    auto kern = t2s_decode_kernel<WT, NB>;
    COVO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    COVO_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T2S_THREADS, smem));
    if (per_sm < 1) return fail(COVO_ERR_INVALID, "t2s decode kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    void* params[1] = {const_cast<T2SDecArgs*>(&args)};
    COVO_CK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(h->di.num_sms), dim3(T2S_THREADS), params,
                                        smem, st));
    return COVO_OK;
}

inline int t2s_sgemm(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, cudaStream_t st) {
    ProfScope ps(PC_PROLOGUE, 2.0 * M * N * K, st);
    sgemm_nt_kernel<float><<<dim3(ceil_div(N, 64), ceil_div(M, 64)), 256, 0, st>>>(A, K, W, bias, C, N, M, N, K, SG_NONE);
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

// Source side of generate (text2semantic.py:716-744) + cross-attention k/v of every decoder layer (:231, :253-260).
inline int t2s_enqueue_source(const covo_t2s* h, const T2SBuffers& t, int NB, int S1, int max_len, cudaStream_t st) {
    const covo_t2s_cfg& c = h->cfg;
    const int M = NB * S1, D = c.dim, inner = h->inner, fi = h->fi_enc, H = c.heads;
    const int rope_n = max_len > S1 ? max_len : S1;
    rope_table_kernel<<<ceil_div(rope_n * (T2S_DH / 2), 256), 256, 0, st>>>(h->inv_freq.as<float>(), t.rope, rope_n, T2S_DH / 2);
    t2s_embed_text_kernel<<<M, 128, 0, st>>>(t.ids, h->enc_emb.as<float>(), t.xe, t.mask, D, c.text_pad_id,
                                             static_cast<long long>(h->enc_emb.shape[0]));
    for (int L = 0; L < c.source_depth; ++L) {
        const T2SEncLayerW& e = h->enc[L];
        t2s_rmsnorm_kernel<<<ceil_div(M, 8), 256, 0, st>>>(t.xe, e.attn_gamma.as<float>(), t.he, M, D);
        COVO_TRY(t2s_sgemm(t.he, e.q_w.as<float>(), nullptr, t.qe, M, inner, D, st));
        COVO_TRY(t2s_sgemm(t.he, e.kv_w.as<float>(), nullptr, t.kve, M, 2 * inner, D, st));
        t2s_rope_rows_kernel<<<ceil_div(M * inner / 2, 256), 256, 0, st>>>(t.qe, t.rope, M, inner, inner, S1);
        t2s_rope_rows_kernel<<<ceil_div(M * inner / 2, 256), 256, 0, st>>>(t.kve, t.rope, M, 2 * inner, inner, S1);
        t2s_enc_attention_kernel<<<NB * H * S1, 128, (S1 + T2S_DH) * sizeof(float), st>>>(t.qe, t.kve, t.mask, t.ae, S1, H);
        COVO_TRY(t2s_sgemm(t.ae, e.out_w.as<float>(), nullptr, t.tmp, M, D, inner, st));
        t2s_add_kernel<<<ceil_div(M * D, 256), 256, 0, st>>>(t.xe, t.tmp, static_cast<size_t>(M) * D);
        t2s_rmsnorm_kernel<<<ceil_div(M, 8), 256, 0, st>>>(t.xe, e.ff_gamma.as<float>(), t.he, M, D);
        COVO_TRY(t2s_sgemm(t.he, e.ff1_w.as<float>(), e.ff1_b.as<float>(), t.f1e, M, 2 * fi, D, st));
        t2s_geglu_kernel<<<ceil_div(M * fi, 256), 256, 0, st>>>(t.f1e, t.ge, M, fi);
        COVO_TRY(t2s_sgemm(t.ge, e.ff2_w.as<float>(), e.ff2_b.as<float>(), t.tmp, M, D, fi, st));
        t2s_add_kernel<<<ceil_div(M * D, 256), 256, 0, st>>>(t.xe, t.tmp, static_cast<size_t>(M) * D);
    }
    t2s_rmsnorm_kernel<<<ceil_div(M, 8), 256, 0, st>>>(t.xe, h->enc_final.as<float>(), t.he, M, D);   // he = source_emb
    const int n_ctx = S1 + 1, n_ctx_r = round_up(n_ctx, 32);
    const size_t per_layer = static_cast<size_t>(NB) * H * n_ctx_r * T2S_DH;
    const int n_scatter = NB * H * n_ctx * (T2S_DH / 2);
    for (int L = 0; L < c.target_depth; ++L) {
        const T2SDecLayerW& d = h->dec[L];
        COVO_TRY(t2s_sgemm(t.he, d.ca_kv.as<float>(), nullptr, t.kve, M, 2 * inner, D, st));
        t2s_ctx_scatter_kernel<<<ceil_div(n_scatter, 256), 256, 0, st>>>(t.kve, d.ca_null.as<float>(), t.mask, t.ctx_k + L * per_layer,
                                                                       t.ctx_v + L * per_layer, t.cmask, NB, S1, H, n_ctx_r,
                                                                       h->wdt == DT_BF16 ? 1 : 0);
    }
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

inline int t2s_source_launches(const covo_t2s* h) { return 2 + 14 * h->cfg.source_depth + 1 + 2 * h->cfg.target_depth; }

}  // namespace covo
