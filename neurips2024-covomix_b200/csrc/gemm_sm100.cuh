// Persistent warp-specialised implicit-GEMM for sm_100a.
//
//   D[z, q, n] = sum_{tap, k} A[z + tap_z[tap], q + tap_row[tap], k] * W[n, tap*Ktap + k]
//
// A: bf16/fp16 activations, 3-D tensor (k contiguous, row, z) read by TMA (out-of-range rows are
// zero-filled by the TMA unit, which is how "same" padding / conv halos / ragged M are handled);
// W: bf16/fp16 weights [N, taps*Ktap] (K-major: exactly nn.Linear's [out, in]; convs are packed
// tap-major by the host).  Accumulation in fp32 in TMEM via tcgen05.mma (cta_group::1,
// M=128 x N=BN x K=16 per instruction), operands staged by TMA into 128B-swizzled smem through an
// mbarrier ring; the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the
// main loop of tile i+1.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..5 = epilogue (TMEM -> registers -> [RoPE] -> smem transpose -> bias / residual /
// activation -> coalesced global stores).
//
// The same kernel serves every dense contraction on the hot path: the velocity net's Linear layers
// (taps = 1), its U-Net skip combiner (taps = 2 over two activation slots), HiFi-GAN's dilated
// Conv1d (taps = kernel size, tap_row = k*dilation - pad) and ConvTranspose1d (polyphase:
// N = stride*Cout, taps = ceil(K/stride), tap_row = -j).
#pragma once
#include "ptx.cuh"

namespace covo {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;                 // 64 x 2 B = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_MAX_TAPS = 16;
constexpr int GEMM_STAGE_A_BYTES = GEMM_BM * GEMM_BK * 2;
constexpr int GEMM_EPI_STAGING_BYTES = 4 * 32 * 64 * 4;   // 4 warps x (32 rows x 64 cols) fp32

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };

struct GemmArgs {
    CUtensorMap tmA;                 // (k, row, z) box (64, 128, 1), SWIZZLE_128B
    CUtensorMap tmB;                 // (k, n)      box (64, BN),     SWIZZLE_128B
    int rows;                        // GEMM rows (q) per z
    int Z;                           // number of z entries (batch items); 1 for Linear layers
    int n_tiles;                     // N_pad / BN
    int n_valid;                     // columns >= n_valid are not stored
    int taps;
    int kc_per_tap;                  // Ktap / 64
    int tap_row[GEMM_MAX_TAPS];
    int tap_z[GEMM_MAX_TAPS];
    // ---- output mapping: element offset = z*out_zs + q*out_rs + n (shared by f32 / 16-bit / residual)
    long long out_zs;
    long long out_rs;
    long long out_off;               // added to the element offset (ConvTranspose: -pad*Cout)
    // row validity: 0 <= up_s*q + n/phase_w - up_p < t_out   (plain GEMM: up_s=1, up_p=0, phase_w=1<<30, t_out=rows)
    int up_s, up_p, phase_w, t_out;
    // ---- epilogue
    const float* bias;               // [N_pad] or null
    const float* residual;           // fp32, same mapping as the output, or null (may alias out_f32)
    float* out_f32;                  // or null
    void* out_h;                     // 16-bit (bf16 or fp16, see h_is_fp16) output or null
    int act_f32;                     // ACT_NONE | ACT_TANH
    int act_h;                       // ACT_NONE | ACT_GELU | ACT_LRELU   (applied to the 16-bit output only)
    float slope;                     // LeakyReLU slope
    int h_is_fp16;                   // 16-bit output format: 0 = bf16, 1 = fp16
    const float2* rope;              // [seq][32] (cos, sin) or null
    int rope_seq;                    // position = q % rope_seq
    int rope_cols;                   // columns < rope_cols are rotated (q and k of to_qkv)
};

template <int BN>
struct GemmCfg {
    static constexpr int STAGE_B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = GEMM_STAGE_A_BYTES + STAGE_B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_EPI_STAGING_BYTES + 256 /*barriers*/ + 1024 /*align*/;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b, int is_fp16) {
    if (is_fp16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN, int AB_FMT /*0 fp16, 1 bf16*/>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs args) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    float* staging = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + GEMM_EPI_STAGING_BYTES);
    uint64_t* full = bars;                         // [STAGES]
    uint64_t* empty = bars + Cfg::STAGES;          // [STAGES]
    uint64_t* tfull = bars + 2 * Cfg::STAGES;      // [2]
    uint64_t* tempty = bars + 2 * Cfg::STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles_per_z = (args.rows + GEMM_BM - 1) / GEMM_BM;
    const int total_tiles = args.Z * m_tiles_per_z * args.n_tiles;
    const int k_iters = args.taps * args.kc_per_tap;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tmA);
        tma_prefetch_desc(&args.tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_idx = tile % args.n_tiles;
                const int mz = tile / args.n_tiles;
                const int m_idx = mz % m_tiles_per_z;
                const int z = mz / m_tiles_per_z;
                const int q0 = m_idx * GEMM_BM;
                const int n0 = n_idx * BN;
                for (int tap = 0; tap < args.taps; ++tap) {
                    const int arow = q0 + args.tap_row[tap];
                    const int az = z + args.tap_z[tap];
                    for (int kc = 0; kc < args.kc_per_tap; ++kc) {
                        mbar_wait(&empty[s], ph ^ 1);
                        uint8_t* sa = stage_base + s * Cfg::STAGE_BYTES;
                        uint8_t* sb = sa + GEMM_STAGE_A_BYTES;
                        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                        tma_load_3d(sa, &args.tmA, &full[s], kc * GEMM_BK, arow, az);
                        tma_load_2d(sb, &args.tmB, &full[s], (tap * args.kc_per_tap + kc) * GEMM_BK, n0);
                        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN, AB_FMT, 0, 0);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_base + s * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + GEMM_STAGE_A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t da = smem_desc_sw128(sa + k * 32, 1024, 16);
                        const uint64_t db = smem_desc_sw128(sb + k * 32, 1024, 16);
                        umma_f16(d_tmem, da, db, idesc, (it | k) != 0);
                    }
                    umma_commit(&empty[s]);
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue warps (4)
        const int lq = warp & 3;                       // TMEM lane quarter this warp may access
        float* stg = staging + (warp - 2) * (32 * 64);
        int as = 0;
        uint32_t aph = 0;
        constexpr int CH = 64;
        constexpr int NCH = BN / CH;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n_idx = tile % args.n_tiles;
            const int mz = tile / args.n_tiles;
            const int m_idx = mz % m_tiles_per_z;
            const int z = mz / m_tiles_per_z;
            const int q_warp0 = m_idx * GEMM_BM + lq * 32;
            const int n0 = n_idx * BN;

            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + static_cast<uint32_t>(as * BN) + (static_cast<uint32_t>(lq * 32) << 16);

#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                uint32_t r[64];
                {
                    uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
                    uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[32]);
                    tmem_ld_32x32(t_acc + c * CH, r0);
                    tmem_ld_32x32(t_acc + c * CH + 32, r1);
                    tmem_ld_wait();
                }
                if (c == NCH - 1) {
                    // accumulator fully read: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                const int ncol0 = n0 + c * CH;
                if (args.rope != nullptr && ncol0 < args.rope_cols) {
                    const int pos = (q_warp0 + lane) % args.rope_seq;
                    const float2* tab = args.rope + static_cast<size_t>(pos) * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 cs = __ldg(tab + j);
                        const float x1 = __uint_as_float(r[j]);
                        const float x2 = __uint_as_float(r[j + 32]);
                        r[j] = __float_as_uint(x1 * cs.x - x2 * cs.y);
                        r[j + 32] = __float_as_uint(x2 * cs.x + x1 * cs.y);
                    }
                }
                // registers (thread = row) -> swizzled smem (conflict-free float4 stores)
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) {
                    float4 v = make_float4(__uint_as_float(r[4 * c4]), __uint_as_float(r[4 * c4 + 1]),
                                           __uint_as_float(r[4 * c4 + 2]), __uint_as_float(r[4 * c4 + 3]));
                    *reinterpret_cast<float4*>(stg + lane * 64 + ((c4 ^ (lane & 7)) << 2)) = v;
                }
                __syncwarp();
                // smem -> global, lane = column pair, coalesced rows
                const int n = ncol0 + 2 * lane;
                const bool n_ok = n < args.n_valid;      // n_valid is even
                float b0 = 0.f, b1 = 0.f;
                if (args.bias != nullptr && n_ok) {
                    b0 = __ldg(args.bias + n);
                    b1 = __ldg(args.bias + n + 1);
                }
                const int phase = n / args.phase_w;
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr) {
                    const int q = q_warp0 + rr;
                    const int t = args.up_s * q + phase - args.up_p;
                    if (!n_ok || t < 0 || t >= args.t_out) continue;
                    const float2 a = *reinterpret_cast<const float2*>(
                        stg + rr * 64 + ((((lane >> 1) ^ (rr & 7)) << 2) | ((lane & 1) << 1)));
                    float v0 = a.x + b0, v1 = a.y + b1;
                    const long long off = static_cast<long long>(z) * args.out_zs + static_cast<long long>(q) * args.out_rs +
                                          n + args.out_off;
                    if (args.residual != nullptr) {
                        const float2 rsd = *reinterpret_cast<const float2*>(args.residual + off);
                        v0 += rsd.x;
                        v1 += rsd.y;
                    }
                    if (args.out_f32 != nullptr) {
                        float o0 = v0, o1 = v1;
                        if (args.act_f32 == ACT_TANH) { o0 = tanhf(o0); o1 = tanhf(o1); }
                        *reinterpret_cast<float2*>(args.out_f32 + off) = make_float2(o0, o1);
                    }
                    if (args.out_h != nullptr) {
                        float o0 = v0, o1 = v1;
                        if (args.act_h == ACT_GELU) { o0 = gelu_erf(o0); o1 = gelu_erf(o1); }
                        else if (args.act_h == ACT_LRELU) { o0 = lrelu(o0, args.slope); o1 = lrelu(o1, args.slope); }
                        *reinterpret_cast<uint32_t*>(static_cast<uint16_t*>(args.out_h) + off) =
                            pack_h2(o0, o1, args.h_is_fp16);
                    }
                }
                __syncwarp();
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace covo
