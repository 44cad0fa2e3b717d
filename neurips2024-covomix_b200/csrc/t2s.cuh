// Text-to-semantic (CoSingle / CoMix) on sm_100a: TextToSemantic.generate (covomix/covomix_model/text2semantic.py:659-848).
//
// Two parts:
//  * the source side (embedding, non-causal source_transformer, cross-attention k/v of every decoder layer) runs once per
//    call as a handful of fp32 kernels -- a few hundred text tokens, < 1 % of the work;
//  * the autoregressive loop is ONE persistent cooperative kernel (t2s_decode_kernel): all SMs walk the stage list of a
//    decoding step together, separated by grid barriers, and loop over steps without returning to the host -- no kernel
//    launch, no host round trip and no re-read of the prefix per token (the reference re-embeds the whole prefix,
//    re-projects the cross-attention context and launches ~200 kernels per step).  A step is a chain of skinny
//    matrix-vector products (B <= 8 rows): it is bound by streaming the decoder weights (46 M parameters, 92 MB as bf16,
//    resident in the 126 MB L2 after the first step) and by the barrier latency, not by the tensor cores, so the stages are
//    CUDA-core GEMVs with 16-byte weight loads, fp32 activations in shared memory and fp32 accumulation.
//    The sampler (top-k filter + Gumbel arg-max, text2semantic.py:104-132) and the EOS logic (:804-826) are stages of
//    the same kernel.
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

// kv [B*S1, 2*inner] of the encoded text + null_kv [2,H,1,64] -> ctx k / v [B][H][1+S1][64] (null first, :253-257);
// cmask [B][1+S1] = 1 | source_mask (:259-260)
__global__ void t2s_ctx_scatter_kernel(const float* __restrict__ kv, const float* __restrict__ null_kv,
                                       const uint8_t* __restrict__ mask, float* __restrict__ ck, float* __restrict__ cv,
                                       uint8_t* __restrict__ cmask, int B, int S1, int H) {
    const int inner = H * T2S_DH;
    const size_t n = static_cast<size_t>(B) * H * (S1 + 1) * T2S_DH;
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = i % T2S_DH;
    const int j = (i / T2S_DH) % (S1 + 1);
    const int h = (i / (static_cast<size_t>(T2S_DH) * (S1 + 1))) % H;
    const int b = i / (static_cast<size_t>(T2S_DH) * (S1 + 1) * H);
    if (j == 0) {
        ck[i] = null_kv[h * T2S_DH + d];
        cv[i] = null_kv[(H + h) * T2S_DH + d];
    } else {
        const float* r = kv + (static_cast<size_t>(b) * S1 + (j - 1)) * 2 * inner + h * T2S_DH + d;
        ck[i] = r[0];
        cv[i] = r[inner];
    }
    if (h == 0 && d == 0) cmask[b * (S1 + 1) + j] = j == 0 ? 1 : mask[b * S1 + j - 1];
}

// ======================================================================================================================
// the persistent decode kernel
// ======================================================================================================================
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
    const float* ctx_k;      // [B][H][n_ctx][64]
    const float* ctx_v;
    float* kcache;           // [B][H][max_len][64]  (rotated keys)
    float* vcache;
};

struct T2SDecArgs {
    T2SLayerW L[T2S_MAX_DEPTH];
    int depth, B, Dt, inner, H, ffi, ffi_pad, n_out, demb, n_logits, n_ctx, max_len, topk, nsplit_self, nsplit_ctx;
    float temperature;
    long long eos_id;
    const float* emb;            // [n_logits, demb] fp32: input embedding and (tied) logit projection
    const float* start;          // [Dt]
    const float* final_gamma;
    const float2* rope;          // [max_len][32]
    const uint8_t* ctx_mask;     // [B][n_ctx]
    float* x;                    // [B][Dt]   residual stream of the current position
    float* q;                    // [B][inner]
    float* part;                 // [B*H*nsplit][66]
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

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs are co-resident (cooperative launch).  A stuck barrier sets the abort flag instead of hanging the device.
__device__ __forceinline__ bool t2s_grid_barrier(const T2SDecArgs& a, unsigned& epoch) {
    __shared__ int s_abort;
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(a.barrier, 1u);
        const long long t0 = clock64();
        int ab = 0;
        while (ld_acquire_u32(a.barrier) < epoch) {
            if (clock64() - t0 > 4000000000ll) {
                a.result[2] = 1;
                ab = 1;
                break;
            }
        }
        if (!ab) ab = *reinterpret_cast<volatile int*>(&a.result[2]);
        s_abort = ab;
        __threadfence();
    }
    __syncthreads();
    return s_abort != 0;
}

template <class WT>
struct WVec;
template <>
struct WVec<__nv_bfloat16> {
    static constexpr int N = 8;     // elements per 16-byte load
    __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&w)[8]) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
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
        const float4 v = *reinterpret_cast<const float4*>(p);
        w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
    }
};

// Skinny product over row PAIRS: for pair p, rows r0 = base(p), r1 = r0 + pair_off of W [*, ldw] against the NB activation
// rows sx[b][0..K) in shared memory; epi(r0, r1, acc0[NB], acc1[NB]) runs on lane 0 with the full sums.  Pairs are dealt
// round-robin to all warps of the grid.  K is a multiple of WVec::N; W rows are 16-byte aligned.
template <class WT, int NB, class Epi>
__device__ __forceinline__ void t2s_gemv_pairs(const WT* __restrict__ W, int ldw, int n_pairs, int pair_stride,
                                               int pair_off, int K, const float* sx, int ldx, Epi epi) {
    constexpr int VN = WVec<WT>::N;
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * T2S_WARPS + (threadIdx.x >> 5);
    const int GW = gridDim.x * T2S_WARPS;
    const int chunks = K / VN;
    for (int p = gw; p < n_pairs; p += GW) {
        const int r0 = p * pair_stride, r1 = r0 + pair_off;
        const WT* w0p = W + static_cast<size_t>(r0) * ldw;
        const WT* w1p = W + static_cast<size_t>(r1) * ldw;
        float acc0[NB], acc1[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc0[b] = acc1[b] = 0.f;
#pragma unroll 4
        for (int c = lane; c < chunks; c += 32) {
            float w0[VN], w1[VN];
            WVec<WT>::load(w0p + c * VN, w0);
            WVec<WT>::load(w1p + c * VN, w1);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float* xb = sx + b * ldx + c * VN;
#pragma unroll
                for (int i = 0; i < VN; i += 4) {
                    const float4 xv = *reinterpret_cast<const float4*>(xb + i);
                    acc0[b] = fmaf(w0[i], xv.x, acc0[b]);
                    acc0[b] = fmaf(w0[i + 1], xv.y, acc0[b]);
                    acc0[b] = fmaf(w0[i + 2], xv.z, acc0[b]);
                    acc0[b] = fmaf(w0[i + 3], xv.w, acc0[b]);
                    acc1[b] = fmaf(w1[i], xv.x, acc1[b]);
                    acc1[b] = fmaf(w1[i + 1], xv.y, acc1[b]);
                    acc1[b] = fmaf(w1[i + 2], xv.z, acc1[b]);
                    acc1[b] = fmaf(w1[i + 3], xv.w, acc1[b]);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc0[b] += __shfl_xor_sync(0xffffffffu, acc0[b], o);
                acc1[b] += __shfl_xor_sync(0xffffffffu, acc1[b], o);
            }
        }
        if (lane == 0) epi(r0, r1, acc0, acc1);
    }
}

// sx[b][0..D) = RMSNorm(x[b]) * gamma  (text2semantic.py:143-151); x is written by other CTAs -> L2 loads (__ldcg)
template <int NB>
__device__ __forceinline__ void t2s_load_norm(const float* x, const float* __restrict__ gamma, int D, float* sx, int ldx,
                                              float* sred) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float ss[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ss[b] = 0.f;
        for (int j = tid; j < D; j += T2S_THREADS) {
            const float v = __ldcg(x + b * D + j);
            sx[b * ldx + j] = v;
            ss[b] = fmaf(v, v, ss[b]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss[b] += __shfl_xor_sync(0xffffffffu, ss[b], o);
        if (lane == 0) sred[b * T2S_WARPS + warp] = ss[b];
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < T2S_WARPS; ++w) t += sred[b * T2S_WARPS + w];
        const float scale = sqrtf(static_cast<float>(D)) / fmaxf(sqrtf(t), 1e-12f);
        for (int j = tid; j < D; j += T2S_THREADS) sx[b * ldx + j] *= scale * gamma[j];
    }
    __syncthreads();
}

// One (b, h, split) unit of single-query attention: q [64] against keys [k0, k1) of K/V [n_alloc][64] -> partial (m, l, o).
// Scores use masked_fill(-FLT_MAX) like the reference (attend_t2s.py:151-153).
__device__ __forceinline__ void t2s_attn_unit(const float* q, const float* K, const float* V, const uint8_t* mask, int k0,
                                              int k1, float* part, float* sq, float* sc, float* sred, float* so) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < T2S_DH) sq[tid] = __ldcg(q + tid) * 0.125f;        // dim_head ** -0.5
    __syncthreads();
    const int nk = k1 - k0;
    float mx = -3.402823466e38f;
    for (int j = tid; j < nk; j += T2S_THREADS) {
        const float4* kr = reinterpret_cast<const float4*>(K + static_cast<size_t>(k0 + j) * T2S_DH);
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < T2S_DH / 4; ++d) {
            const float4 kv = __ldcg(kr + d);
            s = fmaf(sq[4 * d], kv.x, s);
            s = fmaf(sq[4 * d + 1], kv.y, s);
            s = fmaf(sq[4 * d + 2], kv.z, s);
            s = fmaf(sq[4 * d + 3], kv.w, s);
        }
        if (mask != nullptr && !mask[k0 + j]) s = -3.402823466e38f;
        sc[j] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sred[warp] = mx;
    __syncthreads();
    mx = sred[0];
#pragma unroll
    for (int w = 1; w < T2S_WARPS; ++w) mx = fmaxf(mx, sred[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nk; j += T2S_THREADS) {
        const float p = expf(sc[j] - mx);
        sc[j] = p;
        sum += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) sred[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < T2S_WARPS; ++w) sum += sred[w];
    // o = sum_j p_j V[j]: warps stride over keys, a lane owns dims (2 lane, 2 lane + 1) -> 256-byte coalesced rows
    float2 acc = make_float2(0.f, 0.f);
    for (int j = warp; j < nk; j += T2S_WARPS) {
        const float2 vv = __ldcg(reinterpret_cast<const float2*>(V + static_cast<size_t>(k0 + j) * T2S_DH) + lane);
        const float p = sc[j];
        acc.x = fmaf(p, vv.x, acc.x);
        acc.y = fmaf(p, vv.y, acc.y);
    }
    so[warp * T2S_DH + 2 * lane] = acc.x;
    so[warp * T2S_DH + 2 * lane + 1] = acc.y;
    __syncthreads();
    if (tid < T2S_DH) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < T2S_WARPS; ++w) o += so[w * T2S_DH + tid];
        part[2 + tid] = o;
    }
    if (tid == 0) {
        part[0] = nk > 0 ? mx : -3.402823466e38f;
        part[1] = nk > 0 ? sum : 0.f;
    }
    __syncthreads();
}

// sx[b][h*64 + d] = merged attention output of the nsplit partials of (b, h)
template <int NB>
__device__ __forceinline__ void t2s_combine(const float* part, int H, int nsplit, float* sx, int ldx) {
    for (int i = threadIdx.x; i < NB * H * T2S_DH; i += T2S_THREADS) {
        const int d = i % T2S_DH, bh = i / T2S_DH;
        const float* p = part + static_cast<size_t>(bh) * nsplit * T2S_PART;
        float M = -3.402823466e38f;
        for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(p + s * T2S_PART));
        float l = 0.f, o = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float ls = __ldcg(p + s * T2S_PART + 1);
            if (ls > 0.f) {
                const float w = expf(__ldcg(p + s * T2S_PART) - M);
                l = fmaf(ls, w, l);
                o = fmaf(__ldcg(p + s * T2S_PART + 2 + d), w, o);
            }
        }
        sx[(bh / H) * ldx + (bh % H) * T2S_DH + d] = o / l;
    }
    __syncthreads();
}

template <int NB>
__device__ __forceinline__ void t2s_attention_stage(const T2SDecArgs& a, const float* Kb, const float* Vb, int n_alloc,
                                                    int nkeys, int nsplit, const uint8_t* mask, int mask_ld, float* sq,
                                                    float* sc, float* sred, float* so) {
    const int units = NB * a.H * nsplit;
    const int per = (nkeys + nsplit - 1) / nsplit;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int s = u % nsplit, bh = u / nsplit, b = bh / a.H;
        const int k0 = min(s * per, nkeys), k1 = min(k0 + per, nkeys);
        t2s_attn_unit(a.q + static_cast<size_t>(bh) * T2S_DH, Kb + static_cast<size_t>(bh) * n_alloc * T2S_DH,
                      Vb + static_cast<size_t>(bh) * n_alloc * T2S_DH, mask ? mask + b * mask_ld : nullptr, k0, k1,
                      a.part + static_cast<size_t>(u) * T2S_PART, sq, sc, sred, so);
    }
}

// top-k filter + Gumbel arg-max for one (stream, batch row) (text2semantic.py:104-132, :793-800); also feeds the next
// position: x[b][s*demb ...] = emb[token]  (:746-751)
__device__ __forceinline__ void t2s_sample_unit(const T2SDecArgs& a, int s, int b, int step, float* sl, float* sval,
                                                int* sidx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n_logits;
    const size_t off = (static_cast<size_t>(s) * a.B + b) * n;
    for (int r = tid; r < n; r += T2S_THREADS) {
        const float l = __ldcg(a.logits + off + r);
        sl[r] = l;
        if (a.logits_out != nullptr) a.logits_out[static_cast<size_t>(step) * a.n_out * a.B * n + off + r] = l;
    }
    __syncthreads();
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int r = tid; r < n; r += T2S_THREADS) {
        const float l = sl[r];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float o = sl[j];
            rank += (o > l) || (o == l && j < r);
        }
        if (rank < a.topk) {
            const float uu = a.u[static_cast<size_t>(step) * a.n_out * a.B * n + off + r];
            const float g = -logf(fmaxf(-logf(fmaxf(uu, 1e-20f)), 1e-20f));
            const float v = l / fmaxf(a.temperature, 1e-10f) + g;
            if (v > best || (v == best && r < best_i)) best = v, best_i = r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) best = ov, best_i = oi;
    }
    if (lane == 0) sval[warp] = best, sidx[warp] = best_i;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < T2S_WARPS; ++w)
            if (sval[w] > best || (sval[w] == best && sidx[w] < best_i)) best = sval[w], best_i = sidx[w];
        if (best_i >= n) best_i = 0;               // only reachable with non-finite logits
        a.tokens[(static_cast<size_t>(b) * a.n_out + s) * a.max_len + step] = best_i;
        if (best_i == a.eos_id) a.eos_flags[s * a.B + b] = 1;
        sidx[0] = best_i;
    }
    __syncthreads();
    long long fed = sidx[0];
    if (a.forced != nullptr) fed = a.forced[(static_cast<size_t>(b) * a.n_out + s) * a.max_len + step];
    if (fed < 0 || fed >= n) fed = a.eos_id;       // pad (-1) after EOS: never reached for B = 1 (see t2s.py)
    for (int j = tid; j < a.demb; j += T2S_THREADS)
        a.x[static_cast<size_t>(b) * a.Dt + s * a.demb + j] = a.emb[static_cast<size_t>(fed) * a.demb + j];
    __syncthreads();
}

template <class WT, int NB>
__global__ void __launch_bounds__(T2S_THREADS, 1) t2s_decode_kernel(const __grid_constant__ T2SDecArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int ldx = a.ffi_pad > a.Dt ? a.ffi_pad : a.Dt;
    float* sx = smem;                                  // [NB][ldx]
    float* sc = sx + NB * ldx;                         // [max(max_len, n_ctx, n_logits) + 8]
    const int sc_n = ((max(max(a.max_len, a.n_ctx), a.n_logits) + 8) + 3) & ~3;
    float* sq = sc + sc_n;                             // [64]
    float* sred = sq + T2S_DH;                         // [NB * 16]
    float* so = sred + NB * T2S_WARPS;                 // [16][64]
    int* sidx = reinterpret_cast<int*>(so + T2S_WARPS * T2S_DH);   // [16]
    unsigned epoch = 0;
    const int tid = threadIdx.x;
    const WT* dummy = nullptr;
    (void)dummy;

    // position 0 input: the start token (text2semantic.py:746-751)
    for (int i = blockIdx.x * T2S_THREADS + tid; i < NB * a.Dt; i += gridDim.x * T2S_THREADS) a.x[i] = a.start[i % a.Dt];
    if (t2s_grid_barrier(a, epoch)) return;

    for (int step = 0; step < a.max_len; ++step) {
        for (int L = 0; L < a.depth; ++L) {
            const T2SLayerW& w = a.L[L];
            // ---- S1: self-attention q | k | v of the new position; rotary at position `step`; k, v appended to the cache
            t2s_load_norm<NB>(a.x, w.sa_gamma, a.Dt, sx, ldx, sred);
            {
                const float2* rope = a.rope + static_cast<size_t>(step) * (T2S_DH / 2);
                const int inner = a.inner, H = a.H, max_len = a.max_len;
                float* qb = a.q;
                float* kc = w.kcache;
                float* vc = w.vcache;
                t2s_gemv_pairs<WT, NB>(static_cast<const WT*>(w.sa_qkv), a.Dt, 3 * inner / 2, 2, 1, a.Dt, sx, ldx,
                                       [=](int r0, int, const float* a0, const float* a1) {
                                           const int sect = r0 / inner, c = r0 % inner, h = c / T2S_DH, d = c % T2S_DH;
                                           const float2 cs = rope[d >> 1];
#pragma unroll
                                           for (int b = 0; b < NB; ++b) {
                                               float v0 = a0[b], v1 = a1[b];
                                               if (sect < 2) {
                                                   const float t0 = v0 * cs.x - v1 * cs.y;
                                                   v1 = v1 * cs.x + v0 * cs.y;
                                                   v0 = t0;
                                               }
                                               if (sect == 0) {
                                                   qb[b * inner + c] = v0;
                                                   qb[b * inner + c + 1] = v1;
                                               } else {
                                                   float* dst = (sect == 1 ? kc : vc) +
                                                                ((static_cast<size_t>(b) * H + h) * max_len + step) * T2S_DH + d;
                                                   dst[0] = v0;
                                                   dst[1] = v1;
                                               }
                                           }
                                       });
            }
            if (t2s_grid_barrier(a, epoch)) return;
            // ---- S2: causal self-attention of the one query over the step + 1 cached keys
            t2s_attention_stage<NB>(a, w.kcache, w.vcache, a.max_len, step + 1, a.nsplit_self, nullptr, 0, sq, sc, sred, so);
            if (t2s_grid_barrier(a, epoch)) return;
            // ---- S3: to_out + residual
            t2s_combine<NB>(a.part, a.H, a.nsplit_self, sx, ldx);
            {
                float* x = a.x;
                const int Dt = a.Dt;
                t2s_gemv_pairs<WT, NB>(static_cast<const WT*>(w.sa_out), a.inner, a.Dt / 2, 2, 1, a.inner, sx, ldx,
                                       [=](int r0, int r1, const float* a0, const float* a1) {
#pragma unroll
                                           for (int b = 0; b < NB; ++b) {
                                               x[b * Dt + r0] = a0[b] + __ldcg(x + b * Dt + r0);
                                               x[b * Dt + r1] = a1[b] + __ldcg(x + b * Dt + r1);
                                           }
                                       });
            }
            if (t2s_grid_barrier(a, epoch)) return;
            // ---- S4: cross-attention query
            t2s_load_norm<NB>(a.x, w.ca_gamma, a.Dt, sx, ldx, sred);
            {
                float* qb = a.q;
                const int inner = a.inner;
                t2s_gemv_pairs<WT, NB>(static_cast<const WT*>(w.ca_q), a.Dt, a.inner / 2, 2, 1, a.Dt, sx, ldx,
                                       [=](int r0, int r1, const float* a0, const float* a1) {
#pragma unroll
                                           for (int b = 0; b < NB; ++b) {
                                               qb[b * inner + r0] = a0[b];
                                               qb[b * inner + r1] = a1[b];
                                           }
                                       });
            }
            if (t2s_grid_barrier(a, epoch)) return;
            // ---- S5: cross attention over [null kv | encoded text] with the source padding mask
            t2s_attention_stage<NB>(a, w.ctx_k, w.ctx_v, a.n_ctx, a.n_ctx, a.nsplit_ctx, a.ctx_mask, a.n_ctx, sq, sc, sred, so);
            if (t2s_grid_barrier(a, epoch)) return;
            // ---- S6: to_out + residual
            t2s_combine<NB>(a.part, a.H, a.nsplit_ctx, sx, ldx);
            {
                float* x = a.x;
                const int Dt = a.Dt;
                t2s_gemv_pairs<WT, NB>(static_cast<const WT*>(w.ca_out), a.inner, a.Dt / 2, 2, 1, a.inner, sx, ldx,
                                       [=](int r0, int r1, const float* a0, const float* a1) {
#pragma unroll
                                           for (int b = 0; b < NB; ++b) {
                                               x[b * Dt + r0] = a0[b] + __ldcg(x + b * Dt + r0);
                                               x[b * Dt + r1] = a1[b] + __ldcg(x + b * Dt + r1);
                                           }
                                       });
            }
            if (t2s_grid_barrier(a, epoch)) return;
            // ---- S7: FF in-projection + GEGLU: row i (x part) is paired with row ffi + i (gate)
            t2s_load_norm<NB>(a.x, w.ff_gamma, a.Dt, sx, ldx, sred);
            {
                float* hb = a.hbuf;
                const float* b1 = w.ff1_b;
                const int ffi = a.ffi, ffi_pad = a.ffi_pad;
                t2s_gemv_pairs<WT, NB>(static_cast<const WT*>(w.ff1), a.Dt, a.ffi, 1, a.ffi, a.Dt, sx, ldx,
                                       [=](int r0, int r1, const float* a0, const float* a1) {
                                           const float bx = b1[r0], bg = b1[r1];
#pragma unroll
                                           for (int b = 0; b < NB; ++b)
                                               hb[b * ffi_pad + r0] = gelu_erf(a1[b] + bg) * (a0[b] + bx);
                                           (void)ffi;
                                       });
            }
            if (t2s_grid_barrier(a, epoch)) return;
            // ---- S8: FF out-projection + bias + residual
            for (int i = tid; i < NB * a.ffi_pad; i += T2S_THREADS)
                sx[(i / a.ffi_pad) * ldx + i % a.ffi_pad] = __ldcg(a.hbuf + i);
            __syncthreads();
            {
                float* x = a.x;
                const float* b2 = w.ff2_b;
                const int Dt = a.Dt;
                t2s_gemv_pairs<WT, NB>(static_cast<const WT*>(w.ff2), a.ffi_pad, a.Dt / 2, 2, 1, a.ffi_pad, sx, ldx,
                                       [=](int r0, int r1, const float* a0, const float* a1) {
#pragma unroll
                                           for (int b = 0; b < NB; ++b) {
                                               x[b * Dt + r0] = a0[b] + b2[r0] + __ldcg(x + b * Dt + r0);
                                               x[b * Dt + r1] = a1[b] + b2[r1] + __ldcg(x + b * Dt + r1);
                                           }
                                       });
            }
            if (t2s_grid_barrier(a, epoch)) return;
        }
        // ---- S9: final norm + tied logit projection per output stream (text2semantic.py:762-776), fp32 table
        t2s_load_norm<NB>(a.x, a.final_gamma, a.Dt, sx, ldx, sred);
        for (int s = 0; s < a.n_out; ++s) {
            float* lg = a.logits + static_cast<size_t>(s) * NB * a.n_logits;
            const int n = a.n_logits;
            t2s_gemv_pairs<float, NB>(a.emb, a.demb, a.n_logits / 2, 2, 1, a.demb, sx + s * a.demb, ldx,
                                      [=](int r0, int r1, const float* a0, const float* a1) {
#pragma unroll
                                          for (int b = 0; b < NB; ++b) {
                                              lg[b * n + r0] = a0[b];
                                              lg[b * n + r1] = a1[b];
                                          }
                                      });
        }
        if (t2s_grid_barrier(a, epoch)) return;
        // ---- S10: sampling, one CTA per (stream, row); writes the next position's input embedding
        for (int u = blockIdx.x; u < a.n_out * NB; u += gridDim.x)
            t2s_sample_unit(a, u / NB, u % NB, step, sc, sred, sidx);
        if (t2s_grid_barrier(a, epoch)) return;
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
    h->ffi_pad = round_up(h->ffi, 8);
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

inline int t2s_nsplit(int num_sms, int B, int H) {
    int s = num_sms / (B * H);
    return s < 1 ? 1 : (s > 16 ? 16 : s);
}

inline size_t t2s_decode_smem(const covo_t2s* h, int NB, int n_ctx, int max_len) {
    const int ldx = h->ffi_pad > h->cfg.target_transformer_dim ? h->ffi_pad : h->cfg.target_transformer_dim;
    int sc_n = max_len > n_ctx ? max_len : n_ctx;
    if (h->n_logits > sc_n) sc_n = h->n_logits;
    sc_n = (sc_n + 8 + 3) & ~3;
    return sizeof(float) * (static_cast<size_t>(NB) * ldx + sc_n + T2S_DH + NB * T2S_WARPS + T2S_WARPS * T2S_DH + T2S_WARPS);
}

// Workspace layout for (B rows padded to NB, S1 text positions incl. EOS, max_len decode positions)
struct T2SBuffers {
    // source side
    long long* ids;
    uint8_t *mask, *cmask;
    float *xe, *he, *qe, *kve, *ae, *f1e, *ge, *tmp;
    float2* rope;
    // decode
    float *ctx_k, *ctx_v, *kcache, *vcache, *x, *q, *part, *hbuf, *logits;
    int *result, *eos_flags;
    unsigned* barrier;
    size_t zero_from = 0, zero_bytes = 0;     // region cleared before every call
};

inline size_t t2s_layout(const covo_t2s* h, int NB, int S1, int max_len, void* ws, T2SBuffers* b) {
    const covo_t2s_cfg& c = h->cfg;
    Arena a(ws, ~static_cast<size_t>(0));
    const size_t M = static_cast<size_t>(NB) * S1;
    const int n_ctx = S1 + 1, H = c.heads;
    const int rope_n = max_len > S1 ? max_len : S1;
    const int nsplit = t2s_nsplit(h->di.num_sms, NB, H);
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
    const size_t ctx_n = static_cast<size_t>(c.target_depth) * NB * H * n_ctx * T2S_DH;
    const size_t cache_n = static_cast<size_t>(c.target_depth) * NB * H * max_len * T2S_DH;
    t.ctx_k = a.take<float>(ctx_n);
    t.ctx_v = a.take<float>(ctx_n);
    t.kcache = a.take<float>(cache_n);
    t.vcache = a.take<float>(cache_n);
    t.x = a.take<float>(static_cast<size_t>(NB) * c.target_transformer_dim);
    t.q = a.take<float>(static_cast<size_t>(NB) * h->inner);
    t.part = a.take<float>(static_cast<size_t>(NB) * H * nsplit * T2S_PART);
    t.logits = a.take<float>(static_cast<size_t>(h->n_out) * NB * h->n_logits);
    a.off = align_up(a.off, 256);
    t.zero_from = a.off;
    t.hbuf = a.take<float>(static_cast<size_t>(NB) * h->ffi_pad);
    t.result = a.take<int>(4);
    t.eos_flags = a.take<int>(static_cast<size_t>(h->n_out) * NB);
    t.barrier = a.take<unsigned>(4);
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
    sgemm_nt_kernel<<<dim3(ceil_div(N, 64), ceil_div(M, 64)), 256, 0, st>>>(A, W, bias, C, M, N, K, SG_NONE);
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
    const int n_ctx = S1 + 1;
    const size_t per_layer = static_cast<size_t>(NB) * H * n_ctx * T2S_DH;
    for (int L = 0; L < c.target_depth; ++L) {
        const T2SDecLayerW& d = h->dec[L];
        COVO_TRY(t2s_sgemm(t.he, d.ca_kv.as<float>(), nullptr, t.kve, M, 2 * inner, D, st));
        t2s_ctx_scatter_kernel<<<ceil_div(static_cast<int>(per_layer), 256), 256, 0, st>>>(
            t.kve, d.ca_null.as<float>(), t.mask, t.ctx_k + L * per_layer, t.ctx_v + L * per_layer, t.cmask, NB, S1, H);
    }
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

inline int t2s_source_launches(const covo_t2s* h) { return 2 + 14 * h->cfg.source_depth + 1 + 2 * h->cfg.target_depth; }

}  // namespace covo
