// Memory-bound / small kernels of the hot path (everything that is not a dense contraction),
// plus an fp32 SIMT GEMM used for the once-per-call time-conditioning tables.
#pragma once
#include "ptx.cuh"

namespace covo {

// ------------------------------------------------------------------------------------------------
// fp32 SIMT GEMM: C[M,N] = act(A[M,K] * B[N,K]^T + bias[N]).   64x64 tile, 256 threads, 4x4 per thread.
// Used for (a) the time MLP  sinu_pos_emb.1 (acoustic.py:361-365) and (b) all 16 AdaptiveRMSNorm
// to_gamma / to_beta projections (acoustic.py:201) of every evaluation time at once -- they depend
// on t only, so they are tabulated per sample() call instead of being re-evaluated per token batch.
// ------------------------------------------------------------------------------------------------
enum : int { SG_NONE = 0, SG_SILU = 1 };

__device__ __forceinline__ float sg_ldw(const float* p) { return *p; }
__device__ __forceinline__ float sg_ldw(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// C[m, n] = act(sum_k A[m*lda + k] * B[n*K + k] + bias[n]),  C row pitch ldc.  B: fp32 or bf16 (the packed Linear weights).
template <class TB>
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, long long lda, const TB* __restrict__ B,
                                                       const float* __restrict__ bias, float* __restrict__ C, long long ldc,
                                                       int M, int N, int K, int act) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, kk = i & 15;
            As[kk][r] = (m0 + r < M && k0 + kk < K) ? A[static_cast<size_t>(m0 + r) * lda + k0 + kk] : 0.f;
            Bs[kk][r] = (n0 + r < N && k0 + kk < K) ? sg_ldw(B + static_cast<size_t>(n0 + r) * K + k0 + kk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (act == SG_SILU) v = v / (1.0f + expf(-v));
            C[static_cast<size_t>(m) * ldc + n] = v;
        }
    }
}

// LearnedSinusoidalPosEmb (acoustic.py:107-111): out[i, :] = [sin(t_i*w*2pi) | cos(t_i*w*2pi)]
__global__ void time_features_kernel(const float* __restrict__ times, const float* __restrict__ w, float* __restrict__ out,
                                     int n_t, int half) {
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        // same association as the reference: ((t * w) * 2) * pi in fp32
        const float f = times[i] * w[j] * 2.0f * 3.14159265358979323846f;
        out[static_cast<size_t>(i) * 2 * half + j] = sinf(f);
        out[static_cast<size_t>(i) * 2 * half + half + j] = cosf(f);
    }
}

// RotaryEmbedding.forward (acoustic.py:126-130): tab[pos][j] = (cos, sin)(pos * inv_freq[j])
__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float2* __restrict__ tab, int seq, int half) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= seq * half) return;
    const int pos = idx / half, j = idx % half;
    const float f = static_cast<float>(pos) * inv_freq[j];
    tab[idx] = make_float2(cosf(f), sinf(f));
}

// Builds the per-call constant part of the to_embed input (acoustic.py:473-503): rows [0, BN) are the
// conditional branch [emb(ids) | cond], rows [BN, 2BN) the null branch [emb(null id) | null_cond].
// out: bf16 [2*B*N, ldk] (ldk >= S*demb + dim_in, padding columns stay zero).
__global__ void embed_input_kernel(const long long* __restrict__ ids, const float* __restrict__ cond,
                                   const float* __restrict__ table, const float* __restrict__ null_cond,
                                   __nv_bfloat16* __restrict__ out, int BN, int S, int demb, int dim_in, int null_id,
                                   int ldk) {
    const int row = blockIdx.x;              // 0 .. 2*BN-1
    const bool null_branch = row >= BN;
    const int src = null_branch ? row - BN : row;
    __nv_bfloat16* o = out + static_cast<size_t>(row) * ldk;
    for (int s = 0; s < S; ++s) {
        long long id = null_branch ? null_id : ids[static_cast<size_t>(src) * S + s];
        // memory safety only: the host mirror rejects ids outside [0, null_id] like nn.Embedding does (acoustic.py:367-368)
        if (id < 0 || id > null_id) id = null_id;
        const float* e = table + static_cast<size_t>(id) * demb;
        for (int j = threadIdx.x; j < demb; j += blockDim.x) o[s * demb + j] = __float2bfloat16(e[j]);
    }
    const float* c = null_branch ? null_cond : cond + static_cast<size_t>(src) * dim_in;
    for (int j = threadIdx.x; j < dim_in; j += blockDim.x) o[S * demb + j] = __float2bfloat16(c[j]);
    // padding columns are rewritten on every call: they meet zero weight columns, but 0 * NaN != 0 and the
    // workspace belongs to the caller between calls
    for (int j = S * demb + dim_in + threadIdx.x; j < ldk; j += blockDim.x) o[j] = __float2bfloat16(0.f);
}

// ConvPositionEmbed + residual (acoustic.py:153-161, :508): x = gelu(dwconv31(h) + b) + h over time, per channel.
// h, x: fp32 [Bt, N, D]; also writes the bf16 copy of x (the U-Net skip operand of layer 0).
// Each thread: one channel, TB consecutive positions, sliding window in registers.
template <int KS, int TB>
__global__ void __launch_bounds__(256, 2) convpos_kernel(const float* __restrict__ h, const float* __restrict__ wT /*[KS][D]*/,
                                                      const float* __restrict__ bias, float* __restrict__ x,
                                                      __nv_bfloat16* __restrict__ x_h, int N, int D, int reverse) {
    // reverse: last sequence / last positions first (serpentine order after the to_embed GEMM, see flow_enqueue_network)
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = (reverse ? static_cast<int>(gridDim.y) - 1 - static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.y)) * TB;
    const int b = reverse ? static_cast<int>(gridDim.z) - 1 - static_cast<int>(blockIdx.z) : static_cast<int>(blockIdx.z);
    if (c >= D) return;
    const float* hb = h + static_cast<size_t>(b) * N * D + c;
    float w[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) w[k] = wT[k * D + c];
    float win[KS + TB - 1];
#pragma unroll
    for (int i = 0; i < KS + TB - 1; ++i) {
        const int n = n0 + i - KS / 2;
        win[i] = (n >= 0 && n < N) ? hb[static_cast<size_t>(n) * D] : 0.f;
    }
    const float bs = bias[c];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const int n = n0 + t;
        if (n >= N) break;
        float a = bs;
#pragma unroll
        for (int k = 0; k < KS; ++k) a = fmaf(w[k], win[t + k], a);
        const float v = gelu_erf(a) + win[t + KS / 2];
        const size_t off = (static_cast<size_t>(b) * N + n) * D + c;
        x[off] = v;
        x_h[off] = __float2bfloat16(v);
    }
}

// AdaptiveRMSNorm / RMSNorm (acoustic.py:174-175, :198-204): out = x / max(||x||, 1e-12) * sqrt(D) * gamma (+ beta),
// written as the bf16 A operand of the next GEMM.  One warp per row, D = 32 * 4 * V.
// reverse != 0: blocks walk the rows from the last to the first -- the rows the producer kernel wrote LAST are still in L2
// (x is 108 MB at C3, the L2 126 MB: in producer order the consumer would find its first rows already evicted).
template <int V>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                      int M, int reverse) {
    constexpr int D = 128 * V;
    const int blk = reverse ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
    const int row = blk * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    float4 v[V];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i] = xr[i * 32 + lane];
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = sqrtf(static_cast<float>(D)) / fmaxf(sqrtf(ss), 1e-12f);
    uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
        float4 bt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (beta != nullptr) bt = reinterpret_cast<const float4*>(beta)[i * 32 + lane];
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i].x * scale * g.x + bt.x, v[i].y * scale * g.y + bt.y);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i].z * scale * g.z + bt.z, v[i].w * scale * g.w + bt.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        orow[i * 32 + lane] = pk;
    }
}

// Classifier-free-guidance combine (acoustic.py:428) fused with the ODE state update (torchdiffeq
// Euler / Midpoint step) and with the re-quantisation of the next network input:
//   v      = (1 + s) * v_cond - s * v_null                    (s == 0 -> single branch)
//   x_new  = x_base + coef * v
//   if x_out  : x_out  = x_new      (fp32 state; may alias x_base)
//   if v_out  : v_out  = v          (single-evaluation entry point)
//   if xin    : xin[row, :dx] = bf16(x_new) for both CFG branches (rows r and BN + r), ld = ldx
__global__ void cfg_update_kernel(const float* __restrict__ v_pred /*[2*BN or BN, dx]*/, const float* x_base,
                                  float* x_out, float* __restrict__ v_out, __nv_bfloat16* __restrict__ xin, int BN,
                                  int dx, int ldx, float s, float coef, int two_branch) {
    const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= BN * ldx) return;
    const int r = gidx / ldx, c = gidx % ldx;
    if (c >= dx) {                                   // padding columns of the network input: exact zeros, every call
        if (xin != nullptr) {
            xin[static_cast<size_t>(r) * ldx + c] = __float2bfloat16(0.f);
            if (two_branch) xin[static_cast<size_t>(BN + r) * ldx + c] = __float2bfloat16(0.f);
        }
        return;
    }
    const int idx = r * dx + c;
    float v = v_pred[idx];
    if (two_branch) v = (1.0f + s) * v - s * v_pred[static_cast<size_t>(BN) * dx + idx];
    if (v_out != nullptr) v_out[idx] = v;
    if (x_base == nullptr) return;
    const float xn = x_base[idx] + coef * v;
    if (x_out != nullptr) x_out[idx] = xn;
    if (xin != nullptr) {
        const __nv_bfloat16 hb = __float2bfloat16(xn);
        xin[static_cast<size_t>(r) * ldx + c] = hb;
        if (two_branch) xin[static_cast<size_t>(BN + r) * ldx + c] = hb;
    }
}

// x (fp32 [BN, dx]) -> bf16 network input rows for both CFG branches.
__global__ void state_to_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xin, int BN, int dx,
                                      int ldx, int two_branch) {
    const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= BN * ldx) return;
    const int r = gidx / ldx, c = gidx % ldx;
    const __nv_bfloat16 hb = __float2bfloat16(c < dx ? x[r * dx + c] : 0.f);   // padding columns: exact zeros, every call
    xin[static_cast<size_t>(r) * ldx + c] = hb;
    if (two_branch) xin[static_cast<size_t>(BN + r) * ldx + c] = hb;
}

// ------------------------------------------------------------------------------------------------ vocoder
__device__ __forceinline__ uint16_t to_h(float v, int is_fp16) {
    if (is_fp16) {
        __half h = __float2half_rn(v);
        return *reinterpret_cast<uint16_t*>(&h);
    }
    __nv_bfloat16 h = __float2bfloat16(v);
    return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float from_h(uint16_t u, int is_fp16) {
    if (is_fp16) return __half2float(*reinterpret_cast<__half*>(&u));
    return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
}

// mel fp32 [B, C, T] (reference layout) -> 16-bit [B, T, ldc] time-major (padding channels stay zero)
__global__ void mel_to_tc_kernel(const float* __restrict__ mel, uint16_t* __restrict__ out, int C, int T, int ldc,
                                 int is_fp16) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + tx;
        tile[i][tx] = (c < C && t < T) ? mel[(static_cast<size_t>(b) * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + tx;
        // channels [C, ldc) are written as zeros on every call (the workspace belongs to the caller between calls)
        if (t < T && c < ldc) out[(static_cast<size_t>(b) * T + t) * ldc + c] = to_h(tile[tx][i], is_fp16);
    }
}

// Stage output of Generator.forward (models.py:105-112): x = (rb0 + rb1 + rb2) / nk, followed by the
// LeakyReLU that precedes the next ConvTranspose1d (slope 0.1) or conv_post (slope 0.01); 16-bit out.
__global__ void stage_mean_act_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                      uint16_t* __restrict__ out, size_t n4, float inv_nk, float slope, int is_fp16) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 va = reinterpret_cast<const float4*>(a)[i];
    if (b != nullptr) {
        const float4 vb = reinterpret_cast<const float4*>(b)[i];
        va.x += vb.x; va.y += vb.y; va.z += vb.z; va.w += vb.w;
    }
    if (c != nullptr) {
        const float4 vc = reinterpret_cast<const float4*>(c)[i];
        va.x += vc.x; va.y += vc.y; va.z += vc.z; va.w += vc.w;
    }
    ushort4 o;
    o.x = to_h(lrelu(va.x * inv_nk, slope), is_fp16);
    o.y = to_h(lrelu(va.y * inv_nk, slope), is_fp16);
    o.z = to_h(lrelu(va.z * inv_nk, slope), is_fp16);
    o.w = to_h(lrelu(va.w * inv_nk, slope), is_fp16);
    reinterpret_cast<ushort4*>(out)[i] = o;
}

// conv_post + tanh (models.py:112-114): Conv1d(C -> 1, k = 7, pad 3) on the already LeakyReLU'd 16-bit
// activations [B, T, ldc]; one output sample per thread.  out_dtype: 0 = f32, 1 = f16, 2 = i16 (x32768, truncating
// like numpy's astype('int16') in mel_decode_to_wav, monologue_generation.py:55-57).
__global__ void __launch_bounds__(256) conv_post_kernel(const uint16_t* __restrict__ act, const float* __restrict__ w /*[7][ldc]*/,
                                                        const float* __restrict__ bias, void* __restrict__ out, int T, int C, int ldc,
                                                        int is_fp16, int out_dtype) {
    extern __shared__ float w_s[];
    for (int i = threadIdx.x; i < 7 * ldc; i += blockDim.x) w_s[i] = w[i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (t >= T) return;
    float acc = __ldg(bias);
    for (int k = 0; k < 7; ++k) {
        const int tt = t + k - 3;
        if (tt < 0 || tt >= T) continue;
        const uint4* row = reinterpret_cast<const uint4*>(act + (static_cast<size_t>(b) * T + tt) * ldc);
        for (int c8 = 0; c8 < C; c8 += 8) {
            const uint4 u = row[c8 >> 3];
            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc = fmaf(from_h(static_cast<uint16_t>(uu[i] & 0xffff), is_fp16), w_s[k * ldc + c8 + 2 * i], acc);
                acc = fmaf(from_h(static_cast<uint16_t>(uu[i] >> 16), is_fp16), w_s[k * ldc + c8 + 2 * i + 1], acc);
            }
        }
    }
    const float y = tanhf(acc);
    const size_t o = static_cast<size_t>(b) * T + t;
    if (out_dtype == 0) {
        static_cast<float*>(out)[o] = y;
    } else if (out_dtype == 1) {
        static_cast<__half*>(out)[o] = __float2half_rn(y);
    } else {
        static_cast<short*>(out)[o] = static_cast<short>(y * 32768.0f);   // |y| < 1 -> no overflow except y == 1.0
    }
}

// ------------------------------------------------------------------------------------------------ debug / validation
// Naive attention on CUDA cores (one warp per query row); used by the self-tests to validate the
// tcgen05 kernel and as an opt-in debug path (COVO_DEBUG_NAIVE_ATTN=1).  Not a product path.
__global__ void naive_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int N,
                                       int heads, int inner, float scale) {
    const int qpos = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y, seq = blockIdx.z;
    if (qpos >= N) return;
    const size_t ld = 3 * static_cast<size_t>(inner);
    const __nv_bfloat16* base = qkv + static_cast<size_t>(seq) * N * ld;
    const __nv_bfloat16* q = base + qpos * ld + head * 64;
    const float q0 = __bfloat162float(q[lane]), q1 = __bfloat162float(q[lane + 32]);
    float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < N; ++j) {
        const __nv_bfloat16* k = base + j * ld + inner + head * 64;
        const __nv_bfloat16* v = base + j * ld + 2 * inner + head * 64;
        float s = q0 * __bfloat162float(k[lane]) + q1 * __bfloat162float(k[lane + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s *= scale;
        const float mn = fmaxf(m, s);
        const float al = expf(m - mn), p = expf(s - mn);
        l = l * al + p;
        a0 = a0 * al + p * __bfloat162float(v[lane]);
        a1 = a1 * al + p * __bfloat162float(v[lane + 32]);
        m = mn;
    }
    __nv_bfloat16* o = out + (static_cast<size_t>(seq) * N + qpos) * inner + head * 64;
    o[lane] = __float2bfloat16(a0 / l);
    o[lane + 32] = __float2bfloat16(a1 / l);
}

}  // namespace covo
