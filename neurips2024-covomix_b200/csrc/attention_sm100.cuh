// Non-causal, unmasked multi-head attention (dim_head = 64) for sm_100a:  O = softmax(Q K^T / 8) V.
//
// Replaces Attend.forward's einsum/softmax/einsum (covomix/covomix_model/attend.py:110-124), which
// materialises the [B,H,N,N] fp32 score tensor in HBM; here scores never leave the SM.
//
// One CTA = one (sequence b, head h, 128-query tile).  Q/K/V tiles are TMA-loaded straight out of
// the to_qkv GEMM's [B*N, 3*H*64] bf16 output (3-D tensor map: column, position, sequence; rows
// past the end of a sequence are zero-filled by TMA and masked to -inf in the softmax).
//   warp 0   : TMA producer (Q once, then a 2-stage ring of K and V tiles of 128 keys)
//   warp 1   : tcgen05.mma issuer:  S_j = Q K_j^T  (M128 x N128 x K64, both K-major, into TMEM),
//              O_j = P_j V_j (M128 x N64 x K128, A = P from smem, B = V MN-major, into TMEM)
//   warps 2-5: softmax, one query row per thread (TMEM lane == row, so row max / row sum need no
//              shuffles): S_j from TMEM -> running max, exp2, row sum -> P_j as bf16 into 128B-swizzled
//              smem (the A operand of the PV MMA) -> O_{j-1} from TMEM, rescale-and-accumulate in
//              registers.  S, P and O are double-buffered so the MMAs of tile j+1 overlap the softmax
//              of tile j.
#pragma once
#include "ptx.cuh"

namespace covo {

constexpr int ATT_BM = 128;      // queries per CTA
constexpr int ATT_BN = 128;      // keys per iteration
constexpr int ATT_D = 64;        // dim_head
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;                 // 16 KB: one [128 x 64] bf16 tile
constexpr int ATT_P_BYTES = ATT_BM * ATT_BN * 2;             // 32 KB
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES /*Q*/ + 2 * ATT_TILE_BYTES /*K*/ + 2 * ATT_TILE_BYTES /*V*/ +
                               2 * ATT_P_BYTES + 256 + 1024;
constexpr int ATT_TMEM_COLS = 512;                           // S: 2 x 128, O: 2 x 64 (power of two >= 384)

struct AttnArgs {
    CUtensorMap tmQKV;     // (col, pos, seq) over [Bt, N, 3*inner] bf16, box (64, 128, 1), SWIZZLE_128B
    __nv_bfloat16* out;    // [Bt, N, inner]
    int N;                 // sequence length
    int heads;
    int inner;             // heads * 64
    float scale_log2e;     // dim_head^-0.5 * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_TILE_BYTES;            // 2 stages
    uint8_t* sV = sK + 2 * ATT_TILE_BYTES;        // 2 stages
    uint8_t* sP = sV + 2 * ATT_TILE_BYTES;        // 2 buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_P_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2]  MMA -> softmax
    uint64_t* s_free = bars + 11;   // [2]  softmax -> MMA
    uint64_t* p_full = bars + 13;   // [2]  softmax -> MMA
    uint64_t* p_free = bars + 15;   // [2]  MMA -> softmax
    uint64_t* o_full = bars + 17;   // [2]  MMA -> softmax
    uint64_t* o_free = bars + 19;   // [2]  softmax -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x;
    const int head = blockIdx.y;
    const int seq = blockIdx.z;
    const int q0 = q_tile * ATT_BM;
    const int n_kv = (args.N + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tmQKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&p_free[i], 1);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_free[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;            // + buf * 128
    const uint32_t tmem_O = tmem_base + 256;      // + buf * 64

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            mbar_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(sQ, &args.tmQKV, q_full, head * ATT_D, q0, seq);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
                tma_load_3d(sK + st * ATT_TILE_BYTES, &args.tmQKV, &k_full[st], args.inner + head * ATT_D, j * ATT_BN, seq);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
                tma_load_3d(sV + st * ATT_TILE_BYTES, &args.tmQKV, &v_full[st], 2 * args.inner + head * ATT_D, j * ATT_BN,
                            seq);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_f16(ATT_BM, ATT_BN, 1, 0, 0);   // Q K^T : both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(ATT_BM, ATT_D, 1, 0, 1);    // P V   : V is MN-major
            const uint32_t aQ = smem_u32(sQ);
            auto issue_S = [&](int j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_full[st], ph);
                mbar_wait(&s_free[st], ph ^ 1);
                tc_fence_after();
                const uint32_t aK = smem_u32(sK + st * ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) {
                    umma_f16(tmem_S + st * ATT_BN, smem_desc_sw128(aQ + k * 32, 1024, 16),
                             smem_desc_sw128(aK + k * 32, 1024, 16), idesc_s, k != 0);
                }
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[st]);
            };
            mbar_wait(q_full, 0);
            issue_S(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_S(j + 1);
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&v_full[st], ph);
                mbar_wait(&p_full[st], ph);
                mbar_wait(&o_free[st], ph ^ 1);
                tc_fence_after();
                const uint32_t aP = smem_u32(sP + st * ATT_P_BYTES);
                const uint32_t aV = smem_u32(sV + st * ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_BN / 16; ++k) {
                    // A = P: K-major, two 64-key chunks of [128 x 128 B]; B = V: MN-major, 16 keys = 2048 B apart
                    const uint64_t da = smem_desc_sw128(aP + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 1024, 16);
                    const uint64_t db = smem_desc_sw128(aV + k * 2048, 1024, ATT_TILE_BYTES);
                    umma_f16(tmem_O + st * ATT_D, da, db, idesc_o, k != 0);
                }
                umma_commit(&v_empty[st]);
                umma_commit(&p_free[st]);
                umma_commit(&o_full[st]);
            }
        }
    } else {
        // ===================================================== softmax warps: thread == query row
        const int lq = warp & 3;
        const int row = lq * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lq * 32) << 16;
        const float c = args.scale_log2e;
        float m_run = -INFINITY;     // running max of raw scores
        float l_run = 0.f;           // running sum of exp
        float alpha_prev = 0.f;      // rescale factor belonging to the O tile not yet accumulated
        float acc[ATT_D];
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) acc[d] = 0.f;

        auto accumulate_O = [&](int j, float alpha) {
            const int st = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(&o_full[st], ph);
            tc_fence_after();
            uint32_t o[64];
            uint32_t (&o0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[0]);
            uint32_t (&o1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[32]);
            tmem_ld_32x32(tmem_O + st * ATT_D + lane_addr, o0);
            tmem_ld_32x32(tmem_O + st * ATT_D + 32 + lane_addr, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[st]);
#pragma unroll
            for (int d = 0; d < ATT_D; ++d) acc[d] = acc[d] * alpha + __uint_as_float(o[d]);
        };

        for (int j = 0; j < n_kv; ++j) {
            const int st = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(&s_full[st], ph);
            tc_fence_after();
            const int kv_valid = args.N - j * ATT_BN;      // keys >= kv_valid are padding (last tile only)
            // pass 1: row max over the 128 scores of this tile
            float m_tile = -INFINITY;
#pragma unroll 1
            for (int cb = 0; cb < ATT_BN / 32; ++cb) {
                uint32_t s[32];
                tmem_ld_32x32(tmem_S + st * ATT_BN + cb * 32 + lane_addr, s);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = (cb * 32 + i < kv_valid) ? __uint_as_float(s[i]) : -INFINITY;
                    m_tile = fmaxf(m_tile, v);
                }
            }
            const float m_new = fmaxf(m_run, m_tile);
            const float alpha = ex2_approx((m_run - m_new) * c);    // first tile: exp2(-inf) = 0
            const float mc = m_new * c;
            // P buffer must have been consumed by the PV MMA two tiles ago
            mbar_wait(&p_free[st], ph ^ 1);
            // pass 2: p = exp2(s*c - m*c), row sum, bf16 -> swizzled smem
            float l_tile = 0.f;
            uint8_t* prow = sP + st * ATT_P_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll 1
            for (int cb = 0; cb < ATT_BN / 32; ++cb) {
                uint32_t s[32];
                tmem_ld_32x32(tmem_S + st * ATT_BN + cb * 32 + lane_addr, s);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = (cb * 32 + i < kv_valid) ? ex2_approx(__uint_as_float(s[i]) * c - mc) : 0.f;
                    float p1 = (cb * 32 + i + 1 < kv_valid) ? ex2_approx(__uint_as_float(s[i + 1]) * c - mc) : 0.f;
                    __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                    // the row sum uses the rounded probabilities, i.e. exactly what the PV MMA multiplies
                    l_tile += __low2float(h) + __high2float(h);
                    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
                }
                // 32 keys = 64 B = four 16-B pieces; key chunk kc = cb/2, piece index within the 128-B row = (cb&1)*4 + t
                uint8_t* pchunk = prow + (cb >> 1) * ATT_TILE_BYTES;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int piece = ((cb & 1) * 4 + t) ^ (row & 7);
                    *reinterpret_cast<uint4*>(pchunk + piece * 16) =
                        make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
                }
            }
            // S buffer fully read; P written: publish both
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_free[st]);
                mbar_arrive(&p_full[st]);
            }
            l_run = l_run * alpha + l_tile;
            m_run = m_new;
            // deferred accumulation of the previous tile's O (its MMA ran while we did this tile's softmax)
            if (j > 0) accumulate_O(j - 1, alpha_prev);
            alpha_prev = alpha;
        }
        accumulate_O(n_kv - 1, alpha_prev);

        const int qpos = q0 + row;
        if (qpos < args.N) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = args.out + (static_cast<size_t>(seq) * args.N + qpos) * args.inner + head * ATT_D;
#pragma unroll
            for (int d = 0; d < ATT_D; d += 8) {
                uint4 v;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[d] * inv, acc[d + 1] * inv);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(acc[d + 2] * inv, acc[d + 3] * inv);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[d + 4] * inv, acc[d + 5] * inv);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(acc[d + 6] * inv, acc[d + 7] * inv);
                v.x = *reinterpret_cast<uint32_t*>(&h0);
                v.y = *reinterpret_cast<uint32_t*>(&h1);
                v.z = *reinterpret_cast<uint32_t*>(&h2);
                v.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(dst + d) = v;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

}  // namespace covo
