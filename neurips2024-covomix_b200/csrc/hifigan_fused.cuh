// Fused last stage of the HiFi-GAN generator (hifi-gan/models.py:104-114) for narrow stages (<= 32 channels):
//     xs = sum_j ResBlock1_j(x) / num_kernels ;  y = tanh(conv_post(leaky_relu(xs, 0.01)))
// in ONE kernel: the three resblocks (6 dilated convs each), the mean, the final LeakyReLU, conv_post and tanh.
//
// Why: with 31 channels (config_covomix.json: 500 >> 4) and 160 samples per mel frame this stage is 40 % of the vocoder's
// time in layer-by-layer form although it holds 14 % of its FLOPs: every conv round-trips a [B, 160 T, 64-padded] tensor
// through HBM (16-bit operand + fp32 residual), ~18 KB per sample per stage.  Here a CTA owns a tile of 256 output samples
// plus the receptive-field halo (60 samples per side for k = 11, dilations 1/3/5, + 3 for conv_post), keeps the fp32
// residual stream and the 16-bit operands of the tile in shared memory across all 18 convs, and writes only the waveform:
// HBM traffic = the stage input once (128 B per sample, x 1.5 for the halo) + 2-4 B per sample out.
// A conv is a GEMM over (tap, c_in): warp-level mma.sync m16n8k16 (16 samples x 32 c_out per warp step, operands via
// ldmatrix from padded, conflict-free rows; fp32 accumulate).  At 32 channels a tcgen05 tile (M = 128 samples x N = 32) would
// leave the 5th-gen tensor core's N dimension mostly idle and force every intermediate through TMEM -> registers -> shared
// memory in the UMMA canonical layout; the warp-level MMA keeps the conv -> LeakyReLU -> conv chain in registers / smem.
// Zero "same" padding is applied per conv at the true sequence ends (rows outside [0, T) are forced to zero after every
// conv), tile halos inside the sequence are simply recomputed.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace covo {

constexpr int HF_C = 32;            // channels (padded)
constexpr int HF_TP = 256;          // output samples per tile
constexpr int HF_R = 384;           // rows computed per tile (multiple of 16)
constexpr int HF_MARGIN = 32;       // zero rows before / after the operand buffers (>= max tap reach)
constexpr int HF_LDA = 40;          // operand row stride in halves (80 B: ldmatrix conflict-free)
constexpr int HF_LDX = 40;          // fp32 row stride of the residual stream: the conv2 epilogue's 8-byte accesses (lane = (row g, column
                                    // pair q)) hit 32 g + 8 q mod 128 per half-warp -- conflict-free; a 32-float stride put all 8 rows of a
                                    // fragment on the same banks (4-way conflicts on every read-modify-write of the residual)
constexpr int HF_LDS = 33;          // fp32 row stride of the resblock sum (conv_post reads a row per thread)
constexpr int HF_THREADS = 512;
constexpr int HF_MT = 2;             // m-tiles per warp and conv: ceil((HF_R / 16) / (HF_THREADS / 32))
constexpr int HF_MAXK = 11;
constexpr int HF_LDW = HF_MAXK * HF_C + 8;   // weight row stride in halves

constexpr size_t hf_al(size_t x) { return (x + 127) / 128 * 128; }
constexpr size_t HF_OFF_XR = 0;
constexpr size_t HF_OFF_A1 = hf_al(HF_OFF_XR + sizeof(float) * HF_R * HF_LDX);
constexpr size_t HF_OFF_A2 = hf_al(HF_OFF_A1 + sizeof(uint16_t) * (HF_R + 2 * HF_MARGIN) * HF_LDA);
constexpr size_t HF_OFF_SUM = hf_al(HF_OFF_A2 + sizeof(uint16_t) * (HF_R + 2 * HF_MARGIN) * HF_LDA);
constexpr size_t HF_OFF_W1 = hf_al(HF_OFF_SUM + sizeof(float) * (HF_TP + 6) * HF_LDS);
constexpr size_t HF_OFF_W2 = hf_al(HF_OFF_W1 + sizeof(uint16_t) * HF_C * HF_LDW);
constexpr size_t HF_OFF_WP = hf_al(HF_OFF_W2 + sizeof(uint16_t) * HF_C * HF_LDW);
constexpr size_t HF_SMEM_BYTES = hf_al(HF_OFF_WP + sizeof(float) * 7 * HF_C);

struct HifiFusedArgs {
    const float* x;                 // [B, T, ldx] fp32: ConvTranspose1d output of the stage (first 32 channels used)
    void* wav;                      // [B, T]
    const void* w1[3][4];           // convs1 [64, k*64] 16-bit (tap-major, channel-padded), per resblock / dilation
    const void* w2[3][4];
    const float* b1[3][4];
    const float* b2[3][4];
    const float* wpost;             // [7][ld_post] fp32
    const float* bpost;
    int ksize[3], dil[3][4], halo[3];
    int nk, nd, T, ldx, ld_post, H, out_dtype;
    float inv_nk, slope_res, slope_post;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
template <bool FP16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (FP16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    if (FP16) {
        const __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float lrelu_f(float v, float s) { return fmaxf(v, v * s); }      // 0 < s < 1 (checked by the host)

// One Conv1d(32 -> 32, k, dilation d, "same") over rows [mt0*16, mt1*16) of the tile: A = operand rows (with margin),
// W = weights [32][k*32] in smem.  epi(row, n_tile, col, v0, v1) receives the pre-bias sums of two adjacent output channels.
// KT > 0: kernel size known at compile time -- the (tap, 16-channel) steps are fully unrolled and software-pipelined: the
// ldmatrix fragments of step s+1 are requested before the MMAs of step s issue, so a warp keeps the tensor pipe fed on its own
// (with three warps per sub-partition there is little else to hide the ~30-cycle ldmatrix latency).  KT = 0: generic loop.
template <bool FP16, int KT, class Epi>
__device__ __forceinline__ void hf_conv(const uint16_t* A, const uint16_t* W, int k_rt, int d, int mt0, int mt1, Epi epi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = HF_THREADS / 32;
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;      // ldmatrix address roles (A)
    const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;       // (B: two n-tiles per x4)
    const int k = KT > 0 ? KT : k_rt;
    const int half = (k - 1) / 2;
    // a warp owns up to HF_MT consecutive m-tiles so that the weight fragments are reused
    const int per = (mt1 - mt0 + nwarps - 1) / nwarps;
    const int m_begin = mt0 + warp * per, m_end = min(m_begin + per, mt1);
    if (m_begin >= m_end) return;
    float acc[HF_MT][4][4];
#pragma unroll
    for (int i = 0; i < HF_MT; ++i)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[i][n][0] = acc[i][n][1] = acc[i][n][2] = acc[i][n][3] = 0.f;
    if (KT > 0) {
        constexpr int S = 2 * (KT > 0 ? KT : 1);                  // steps: (tap, 16-channel half)
        static_assert(HF_MT == 2, "the pipelined path is written for up to two m-tiles per warp");
        const bool two = m_begin + 1 < m_end;                      // warp-uniform: does this warp own a second m-tile?
        const uint16_t* wp0 = W + b_row * HF_LDW + b_col;
        const uint16_t* ap0 = A + (HF_MARGIN + m_begin * 16 - half * d + a_row) * HF_LDA + a_col;
        const int dstep = d * HF_LDA;
        // the warp's m-tile count (1 or 2) is warp-uniform: two straight-line copies instead of a per-step predicate
        // around the .sync.aligned instructions (which ptxas guards with WARPSYNC / BSSY pairs)
        auto run = [&](auto mt_c) {
            constexpr int MT = decltype(mt_c)::value;
            uint32_t bq[2][8], aq[2][MT][4];
            auto fetch = [&](int s, uint32_t (&bb)[8], uint32_t (&aa)[MT][4]) {
                const int tap = s >> 1, ks = s & 1;
                uint32_t (&b01)[4] = *reinterpret_cast<uint32_t (*)[4]>(&bb[0]);
                uint32_t (&b23)[4] = *reinterpret_cast<uint32_t (*)[4]>(&bb[4]);
                ldmatrix_x4(b01, wp0 + tap * HF_C + ks * 16);
                ldmatrix_x4(b23, wp0 + tap * HF_C + ks * 16 + 16 * HF_LDW);
                const uint16_t* ap = ap0 + tap * dstep + ks * 16;
#pragma unroll
                for (int i = 0; i < MT; ++i) ldmatrix_x4(aa[i], ap + i * 16 * HF_LDA);
            };
            fetch(0, bq[0], aq[0]);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int cur = s & 1;
                if (s + 1 < S) fetch(s + 1, bq[cur ^ 1], aq[cur ^ 1]);
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    mma_16816<FP16>(acc[i][0], aq[cur][i], bq[cur][0], bq[cur][1]);
                    mma_16816<FP16>(acc[i][1], aq[cur][i], bq[cur][2], bq[cur][3]);
                    mma_16816<FP16>(acc[i][2], aq[cur][i], bq[cur][4], bq[cur][5]);
                    mma_16816<FP16>(acc[i][3], aq[cur][i], bq[cur][6], bq[cur][7]);
                }
            }
        };
        if (two) run(std::integral_constant<int, 2>());
        else run(std::integral_constant<int, 1>());
    } else {
        for (int tap = 0; tap < k; ++tap) {
            const int shift = (tap - half) * d;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t b01[4], b23[4];
                const uint16_t* wp = W + b_row * HF_LDW + tap * HF_C + ks * 16 + b_col;
                ldmatrix_x4(b01, wp);                      // n-tiles 0, 1
                ldmatrix_x4(b23, wp + 16 * HF_LDW);        // n-tiles 2, 3
#pragma unroll
                for (int i = 0; i < HF_MT; ++i) {
                    if (m_begin + i < m_end) {
                        uint32_t a[4];
                        ldmatrix_x4(a, A + (HF_MARGIN + (m_begin + i) * 16 + shift + a_row) * HF_LDA + ks * 16 + a_col);
                        mma_16816<FP16>(acc[i][0], a, b01[0], b01[1]);
                        mma_16816<FP16>(acc[i][1], a, b01[2], b01[3]);
                        mma_16816<FP16>(acc[i][2], a, b23[0], b23[1]);
                        mma_16816<FP16>(acc[i][3], a, b23[2], b23[3]);
                    }
                }
            }
        }
    }
    const int g = lane >> 2, c2 = (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < HF_MT; ++i) {
        if (m_begin + i < m_end) {
            const int r = (m_begin + i) * 16 + g;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                epi(r, n, n * 8 + c2, acc[i][n][0], acc[i][n][1]);
                epi(r + 8, n, n * 8 + c2, acc[i][n][2], acc[i][n][3]);
            }
        }
    }
}

template <bool FP16>
__device__ __forceinline__ void hf_load_weights(uint16_t* Ws, const void* wg, int k) {
    // global: [64 c_out][k][64 c_in] 16-bit  ->  smem [32][k*32] (row stride HF_LDW), 16-byte pieces
    const uint16_t* g = static_cast<const uint16_t*>(wg);
    const int pieces = HF_C * k * (HF_C / 8);
    for (int i = threadIdx.x; i < pieces; i += HF_THREADS) {
        const int c8 = i % (HF_C / 8), tap = (i / (HF_C / 8)) % k, co = i / ((HF_C / 8) * k);
        const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(Ws + co * HF_LDW + tap * HF_C + c8 * 8));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + (static_cast<size_t>(co) * k + tap) * 64 + c8 * 8));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");      // asynchronous: the caller waits with hf_wait_weights
}
__device__ __forceinline__ void hf_wait_weights() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The dilated-conv chain of resblock j over tile rows [mt0*16, mt1*16) (models.py:35-42); W1 must already be in flight.
// INTERIOR: no row of the tile lies outside the sequence (CTA-uniform, decided once per tile) -- the per-element sequence-end
// tests vanish from the epilogues at compile time (as uniform run-time branches they cost ~25 cycles each, 16 per conv and lane).
template <bool FP16, int KT, bool INTERIOR>
__device__ __forceinline__ void hf_resblock(const HifiFusedArgs& a, int j, int mt0, int mt1, int base, float* XR, uint16_t* A1,
                                            uint16_t* A2, uint16_t* W1, uint16_t* W2) {
    const int k = a.ksize[j];
    auto in_seq = [&](int r) { const int gp = base + r; return INTERIOR || (gp >= 0 && gp < a.T); };
    for (int m = 0; m < a.nd; ++m) {
        hf_wait_weights();
        __syncthreads();                              // W1 landed; A1 / XR of the previous step complete
        hf_load_weights<FP16>(W2, a.w2[j][m], k);
        // xt = lrelu(c1(lrelu(x)))                                  (models.py:36-38)
        {
            float2 bias[4];                                   // this lane's 4 column pairs (c = 8 n + 2 (lane & 3))
#pragma unroll
            for (int n = 0; n < 4; ++n) bias[n] = *reinterpret_cast<const float2*>(a.b1[j][m] + n * 8 + (threadIdx.x & 3) * 2);
            const float slope = a.slope_res;
            hf_conv<FP16, KT>(A1, W1, k, a.dil[j][m], mt0, mt1, [&](int r, int n, int c, float v0, float v1) {
                uint32_t pk = 0u;
                if (in_seq(r)) pk = pack_h2<FP16>(lrelu_f(v0 + bias[n].x, slope), lrelu_f(v1 + bias[n].y, slope));
                *reinterpret_cast<uint32_t*>(A2 + (HF_MARGIN + r) * HF_LDA + c) = pk;
            });
        }
        hf_wait_weights();
        __syncthreads();                              // W2 landed; A2 complete; W1 free
        if (m + 1 < a.nd) hf_load_weights<FP16>(W1, a.w1[j][m + 1], k);
        // x = c2(xt) + x ; next operand lrelu(x)                    (models.py:39-41)
        {
            float2 bias[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) bias[n] = *reinterpret_cast<const float2*>(a.b2[j][m] + n * 8 + (threadIdx.x & 3) * 2);
            const float slope = a.slope_res;
            hf_conv<FP16, KT>(A2, W2, k, 1, mt0, mt1, [&](int r, int n, int c, float v0, float v1) {
                float2 xv = make_float2(0.f, 0.f);
                if (in_seq(r)) {
                    const float2 old = *reinterpret_cast<const float2*>(XR + r * HF_LDX + c);
                    xv = make_float2(v0 + bias[n].x + old.x, v1 + bias[n].y + old.y);
                }
                *reinterpret_cast<float2*>(XR + r * HF_LDX + c) = xv;
                *reinterpret_cast<uint32_t*>(A1 + (HF_MARGIN + r) * HF_LDA + c) =
                    pack_h2<FP16>(lrelu_f(xv.x, slope), lrelu_f(xv.y, slope));
            });
        }
    }
}

template <bool FP16>
__global__ void __launch_bounds__(HF_THREADS, 1) hifigan_fused_last_stage_kernel(const HifiFusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* XR = reinterpret_cast<float*>(smem_raw + HF_OFF_XR);          // [HF_R][HF_LDX] fp32 residual stream
    uint16_t* A1 = reinterpret_cast<uint16_t*>(smem_raw + HF_OFF_A1);    // [HF_R + 2*margin][HF_LDA] lrelu(x)
    uint16_t* A2 = reinterpret_cast<uint16_t*>(smem_raw + HF_OFF_A2);    //  "  lrelu(conv1)
    float* SUM = reinterpret_cast<float*>(smem_raw + HF_OFF_SUM);        // [HF_TP + 6][HF_LDS]
    uint16_t* W1 = reinterpret_cast<uint16_t*>(smem_raw + HF_OFF_W1);    // [32][HF_LDW]
    uint16_t* W2 = reinterpret_cast<uint16_t*>(smem_raw + HF_OFF_W2);
    float* WP = reinterpret_cast<float*>(smem_raw + HF_OFF_WP);          // [7][32] conv_post weights

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * HF_TP;                    // first output sample of the tile
    const int base = p0 - 3 - a.H;                        // global sample index of tile row 0
    const float* xg = a.x + static_cast<size_t>(b) * a.T * a.ldx;
    auto in_seq = [&](int r) { const int gp = base + r; return gp >= 0 && gp < a.T; };

    // zero the operand buffers (margins stay zero for good) and the sum, stage the conv_post weights
    for (int i = tid; i < (HF_R + 2 * HF_MARGIN) * HF_LDA / 8; i += HF_THREADS) {
        reinterpret_cast<uint4*>(A1)[i] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(A2)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < (HF_TP + 6) * HF_LDS; i += HF_THREADS) SUM[i] = 0.f;
    for (int i = tid; i < 7 * HF_C; i += HF_THREADS) WP[i] = a.wpost[(i / HF_C) * a.ld_post + i % HF_C];

    for (int j = 0; j < a.nk; ++j) {
        const int k = a.ksize[j];
        // rows this resblock needs: the TP + 6 output rows plus its own halo, rounded out to whole m-tiles
        const int r_lo = a.H - a.halo[j], r_hi = a.H + HF_TP + 6 + a.halo[j];
        const int mt0 = r_lo / 16, mt1 = (r_hi + 15) / 16;
        // (re)load the stage input (L2-resident after the first resblock): XR = x, A1 = lrelu(x); zero outside the sequence
        __syncthreads();
        hf_load_weights<FP16>(W1, a.w1[j][0], k);         // cp.async: lands while the input tile is staged
        for (int i = tid; i < HF_R * (HF_C / 4); i += HF_THREADS) {
            const int r = i / (HF_C / 4), c4 = (i % (HF_C / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in_seq(r)) v = *reinterpret_cast<const float4*>(xg + static_cast<size_t>(base + r) * a.ldx + c4);
            *reinterpret_cast<float4*>(XR + r * HF_LDX + c4) = v;
            uint2 pk;
            pk.x = pack_h2<FP16>(lrelu_f(v.x, a.slope_res), lrelu_f(v.y, a.slope_res));
            pk.y = pack_h2<FP16>(lrelu_f(v.z, a.slope_res), lrelu_f(v.w, a.slope_res));
            *reinterpret_cast<uint2*>(A1 + (HF_MARGIN + r) * HF_LDA + c4) = pk;
        }
        // weights stream through two buffers with cp.async: c2's arrive while c1 runs, the next c1's while c2 runs
        const bool interior = base >= 0 && base + HF_R <= a.T;
#define HF_RB(KT_)                                                                                   \
    do {                                                                                             \
        if (interior) hf_resblock<FP16, KT_, true>(a, j, mt0, mt1, base, XR, A1, A2, W1, W2);        \
        else hf_resblock<FP16, KT_, false>(a, j, mt0, mt1, base, XR, A1, A2, W1, W2);                \
    } while (0)
        switch (k) {
            case 3: HF_RB(3); break;
            case 5: HF_RB(5); break;
            case 7: HF_RB(7); break;
            case 11: HF_RB(11); break;
            default: HF_RB(0); break;
        }
#undef HF_RB
        __syncthreads();
        // xs += resblock output over the TP + 6 rows conv_post needs
        for (int i = tid; i < (HF_TP + 6) * HF_C; i += HF_THREADS) {
            const int r = i / HF_C, c = i % HF_C;
            SUM[r * HF_LDS + c] += XR[(a.H + r) * HF_LDX + c];
        }
    }
    __syncthreads();
    // y = tanh(conv_post(leaky_relu(xs / nk, 0.01)))                     (models.py:111-114); one sample per thread
    for (int i = tid; i < (HF_TP + 6) * HF_C; i += HF_THREADS) {
        const int r = i / HF_C, c = i % HF_C;
        SUM[r * HF_LDS + c] = lrelu_f(SUM[r * HF_LDS + c] * a.inv_nk, a.slope_post);
    }
    __syncthreads();
    for (int t = tid; t < HF_TP; t += HF_THREADS) {
        const int gp = p0 + t;
        if (gp >= a.T) break;
        float acc = a.bpost[0];
#pragma unroll
        for (int kk = 0; kk < 7; ++kk) {
            const float* row = SUM + (t + kk) * HF_LDS;
#pragma unroll
            for (int c = 0; c < HF_C; ++c) acc = fmaf(row[c], WP[kk * HF_C + c], acc);
        }
        const float y = tanhf(acc);
        const size_t o = static_cast<size_t>(b) * a.T + gp;
        if (a.out_dtype == 0) static_cast<float*>(a.wav)[o] = y;
        else if (a.out_dtype == 1) static_cast<__half*>(a.wav)[o] = __float2half_rn(y);
        else static_cast<short*>(a.wav)[o] = static_cast<short>(y * 32768.0f);
    }
}

}  // namespace covo
