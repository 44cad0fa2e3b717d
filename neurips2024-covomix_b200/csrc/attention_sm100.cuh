// Non-causal, unmasked multi-head attention (dim_head = 64) for sm_100a:  O = softmax(Q K^T / 8) V.
//
// Replaces Attend.forward's einsum/softmax/einsum (covomix/covomix_model/attend.py:110-124), which
// materialises the [B,H,N,N] fp32 score tensor in HBM; here scores never leave the SM.
//
// Persistent: grid = min(#work items, #SMs); a work item = (sequence b, head h, 256-query tile), and the TMA / MMA /
// softmax pipelines run straight across item boundaries (Q is double-buffered), so TMEM allocation, barrier set-up,
// the first Q/K loads and the output stores of one item hide behind the neighbouring items' work.
// A work item = two 128-query groups A and B that share every
// K/V tile and ping-pong on the tensor pipe: while the softmax warps of one group work on S_j, the
// MMAs of the other group run.  Q/K/V tiles are TMA-loaded straight out of the to_qkv GEMM's
// [B*N, 3*H*64] bf16 output (3-D tensor map: column, position, sequence; rows past the end of a
// sequence are zero-filled by TMA and masked to -inf in the softmax).
//   warp 0      : TMA producer (both Q tiles once, then a 2-stage ring of K and V tiles of 128 keys)
//   warp 1      : tcgen05.mma issuer:  S_g = Q_g K_j^T (M128 x N128 x K64, K-major operands, into TMEM),
//                 O_g = P_g V_j (M128 x N64 x K128, A = P from TMEM, B = V MN-major, into TMEM);
//                 S_g(j) is issued before P_g(j-1) V_{j-1} so the pipe always has work queued
//   warps 4-7   : softmax group A, warps 8-11: group B; one query row per thread (TMEM lane == row, so
//                 row max / row sum need no shuffles): S_j from TMEM into registers (buffer released at
//                 once) -> running max, exp2, row sum -> P_j as bf16 pairs into TMEM (the A
//                 operand of the PV MMA; written with tcgen05.st, so no smem round trip and no proxy fence) -> O_{j-1} from TMEM, rescale-and-accumulate in registers.
// Registers are rebalanced with setmaxnreg (producer/MMA warpgroup 56, softmax warpgroups 224; the pool is exactly what warpgroup 0 releases).
#pragma once
#include "ptx.cuh"

namespace covo {

// Optional clock64 timeline of the softmax loop (tools/micro/attn_trace.cu); compiled out in the library.
#ifdef ATT_TRACE
__device__ long long g_trace[4096];
#define TR(slot) do { if (blockIdx.x == 0 && lane == 0 && it >= 40 && it < 44) g_trace[(warp) * 256 + (it - 40) * 16 + (slot)] = clock64(); } while (0)
#else
#define TR(slot)
#endif

constexpr int ATT_BM = 256;      // queries per CTA (two groups of 128)
constexpr int ATT_BN = 128;      // keys per iteration
constexpr int ATT_D = 64;        // dim_head
constexpr int ATT_THREADS = 384; // warpgroup 0: TMA + MMA (+2 idle warps); warpgroups 1, 2: softmax A, B
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;                 // 16 KB: one [128 x 64] bf16 tile
constexpr int ATT_SMEM_BYTES = 4 * ATT_TILE_BYTES /*Q x2 buffers*/ + 2 * ATT_TILE_BYTES /*K*/ + 2 * ATT_TILE_BYTES /*V*/ +
                               256 + 1024;
constexpr int ATT_TMEM_COLS = 512;                           // S: 2 x 128, O: 2 x 64, P (bf16 pairs): 2 x 64

struct AttnArgs {
    CUtensorMap tmQKV;     // (col, pos, seq) over [Bt, N, 3*inner] bf16, box (64, 128, 1), SWIZZLE_128B
    __nv_bfloat16* out;    // [Bt, N, inner]
    int N;                 // sequence length
    int heads;
    int inner;             // heads * 64
    float scale_log2e;     // dim_head^-0.5 * log2(e)
    int n_qt;              // query tiles per (sequence, head) = ceil(N / 256)
    int n_items;           // n_qt * heads * sequences
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {   // sm_100 three-input max: one FMNMX3 instead of two FMNMX
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__global__ void __launch_bounds__(ATT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                           // 2 buffers x 2 groups
    uint8_t* sK = sQ + 4 * ATT_TILE_BYTES;        // 2 stages
    uint8_t* sV = sK + 2 * ATT_TILE_BYTES;        // 2 stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * ATT_TILE_BYTES);
    uint64_t* q_full = bars + 0;    // [2 buffers]   (bars + 0, bars + 23)
    uint64_t* q_empty = bars + 21;  // [2 buffers]
    uint64_t* k_full = bars + 1;    // [2 stages]
    uint64_t* k_empty = bars + 3;
    uint64_t* v_full = bars + 5;
    uint64_t* v_empty = bars + 7;
    uint64_t* s_full = bars + 9;    // [2 groups]  MMA -> softmax
    uint64_t* s_free = bars + 11;   //             softmax -> MMA
    uint64_t* p_full = bars + 13;   //             softmax -> MMA
    uint64_t* o_full = bars + 17;   //             MMA -> softmax
    uint64_t* o_free = bars + 19;   //             softmax -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kv = (args.N + ATT_BN - 1) / ATT_BN;
    uint64_t* q_full1 = bars + 23;  // second Q buffer's full barrier

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tmQKV);
        mbar_init(q_full, 1);
        mbar_init(q_full1, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_empty[i], 1);
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
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
    const uint32_t tmem_S = tmem_base;            // + group * 128
    const uint32_t tmem_O = tmem_base + 256;      // + group * 64
    const uint32_t tmem_P = tmem_base + 384;      // + group * 64: P as bf16 pairs, the A operand of the PV MMA

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0 && elect_one()) {
            // ===================================================== TMA producer
            uint32_t it = 0;                                   // key-tile counter across work items
            int il = 0;                                        // local work-item counter
            for (int w = blockIdx.x; w < args.n_items; w += gridDim.x, ++il) {
                const int q0 = (w % args.n_qt) * ATT_BM;
                const int head = (w / args.n_qt) % args.heads;
                const int seq = w / (args.n_qt * args.heads);
                const int qb = il & 1;
                uint64_t* qf = qb ? q_full1 : q_full;
                mbar_wait(&q_empty[qb], ((il >> 1) & 1) ^ 1);
                mbar_expect_tx(qf, 2 * ATT_TILE_BYTES);
                tma_load_3d(sQ + (2 * qb) * ATT_TILE_BYTES, &args.tmQKV, qf, head * ATT_D, q0, seq);
                tma_load_3d(sQ + (2 * qb + 1) * ATT_TILE_BYTES, &args.tmQKV, qf, head * ATT_D, q0 + 128, seq);
                for (int j = 0; j < n_kv; ++j, ++it) {
                    const int st = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    mbar_wait(&k_empty[st], ph ^ 1);
                    mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
                    tma_load_3d(sK + st * ATT_TILE_BYTES, &args.tmQKV, &k_full[st], args.inner + head * ATT_D, j * ATT_BN, seq);
                    mbar_wait(&v_empty[st], ph ^ 1);
                    mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
                    tma_load_3d(sV + st * ATT_TILE_BYTES, &args.tmQKV, &v_full[st], 2 * args.inner + head * ATT_D, j * ATT_BN,
                                seq);
                }
            }
        } else if (warp == 1 && elect_one()) {
            // ===================================================== MMA issuer (elect.sync: ptxas keeps operands in uniform registers)
            constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_BN, 1, 0, 0);   // Q K^T : both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 1, 0, 1);    // P V   : V is MN-major
            // Descriptors are built once; per MMA only the 14-bit start-address field is advanced (one add), so the
            // single issuing thread never becomes the bottleneck (24 MMAs per key tile).
            uint64_t dQ[4], dK[2], dV[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                dK[i] = smem_desc_sw128(smem_u32(sK + i * ATT_TILE_BYTES), 1024, 16);
                dV[i] = smem_desc_sw128(smem_u32(sV + i * ATT_TILE_BYTES), 1024, ATT_TILE_BYTES);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) dQ[i] = smem_desc_sw128(smem_u32(sQ + i * ATT_TILE_BYTES), 1024, 16);
            // One software pipeline over ALL key tiles of ALL work items of this CTA: at step t issue S(t), then P(t-1) V(t-1).
            int n_my = 0;
            for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) ++n_my;
            const uint32_t total = static_cast<uint32_t>(n_my) * n_kv;
            int il = 0, j = 0;                                 // work item / key tile of step t
            for (uint32_t t = 0; t <= total; ++t) {
                if (t < total) {
                    const int qb = il & 1;
                    if (j == 0) mbar_wait(qb ? q_full1 : q_full, (il >> 1) & 1);
                    const int st = t & 1;
                    mbar_wait(&k_full[st], (t >> 1) & 1);
                    const uint64_t bK = st ? dK[1] : dK[0];
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        mbar_wait(&s_free[g], (t & 1) ^ 1);        // softmax has pulled S_g(t-1) into registers
                        tc_fence_after();
                        const uint64_t bQ = qb ? dQ[2 + g] : dQ[g];
#pragma unroll
                        for (int k = 0; k < ATT_D / 16; ++k)
                            umma_f16(tmem_S + g * ATT_BN, desc_advance(bQ, k * 32), desc_advance(bK, k * 32), idesc_s, k != 0);
                        umma_commit(&s_full[g]);
                    }
                    umma_commit(&k_empty[st]);
                    if (j == n_kv - 1) umma_commit(&q_empty[qb]);  // last S of this item: its Q buffer may be refilled
                    if (++j == n_kv) { j = 0; ++il; }
                }
                if (t >= 1) {
                    const uint32_t tt = t - 1;
                    const int st = tt & 1;
                    mbar_wait(&v_full[st], (tt >> 1) & 1);
                    const uint64_t bV = st ? dV[1] : dV[0];
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        mbar_wait(&p_full[g], tt & 1);
                        mbar_wait(&o_free[g], (tt & 1) ^ 1);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < ATT_BN / 16; ++k) {
                            // A = P from TMEM (16 keys = 8 columns per step); B = V: MN-major, 16 keys = 2048 B apart
                            umma_f16_ts(tmem_O + g * ATT_D, tmem_P + g * 64 + k * 8, desc_advance(bV, k * 2048), idesc_o, k != 0);
                        }
                        umma_commit(&o_full[g]);                   // also means: P_g has been consumed
                    }
                    umma_commit(&v_empty[st]);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===================================================== softmax warpgroups: thread == query row
        const int g = (warp - 4) >> 2;                 // 0: group A (warps 4-7), 1: group B (warps 8-11)
        const int lq = warp & 3;                       // TMEM lane quarter
        const int row = lq * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lq * 32) << 16;
        const uint32_t tS = tmem_S + g * ATT_BN + lane_addr;
        const uint32_t tO = tmem_O + g * ATT_D + lane_addr;
        const float c = args.scale_log2e;
        float acc[ATT_D];
        uint32_t it = 0;             // key-tile counter across work items (barrier parities)
        const uint32_t tP = tmem_P + g * 64 + lane_addr;

        auto accumulate_O = [&](uint32_t t, float alpha) {
            mbar_wait(&o_full[g], t & 1);
            tc_fence_after();
            uint32_t o[64];
            uint32_t (&o0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[0]);
            uint32_t (&o1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[32]);
            tmem_ld_32x32(tO, o0);
            tmem_ld_32x32(tO + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[g]);
#pragma unroll
            for (int d = 0; d < ATT_D; ++d) acc[d] = fmaf(acc[d], alpha, __uint_as_float(o[d]));
        };

        for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) {
        const int q0 = (w % args.n_qt) * ATT_BM;
        const int head = (w / args.n_qt) % args.heads;
        const int seq = w / (args.n_qt * args.heads);
        float m_run = -INFINITY;     // running max of raw scores
        float l_run = 0.f;           // running sum of exp
        float alpha_prev = 0.f;      // rescale factor belonging to the O tile not yet accumulated
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) acc[d] = 0.f;
        for (int j = 0; j < n_kv; ++j, ++it) {
            TR(0);
            mbar_wait(&s_full[g], it & 1);
            TR(1);
            tc_fence_after();
            uint32_t s[128];
            {
                uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
                uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
                uint32_t (&s2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[64]);
                uint32_t (&s3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[96]);
                tmem_ld_32x32(tS, s0);
                tmem_ld_32x32(tS + 32, s1);
                tmem_ld_32x32(tS + 64, s2);
                tmem_ld_32x32(tS + 96, s3);
                tmem_ld_wait();
            }
            // S is in registers: the MMA warp may overwrite the buffer with S(j+1)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
            TR(2);

            const int kv_valid = args.N - j * ATT_BN;      // keys >= kv_valid are padding (last tile only)
            if (kv_valid < ATT_BN) {
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
            }
            // row max: 8 independent chains of three-input max (dependency depth 8 instead of 32)
            float mx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) mx[i] = __uint_as_float(s[i]);
#pragma unroll
            for (int i = 8; i + 15 < 128; i += 16) {
#pragma unroll
                for (int q = 0; q < 8; ++q) mx[q] = fmax3(mx[q], __uint_as_float(s[i + 2 * q]), __uint_as_float(s[i + 2 * q + 1]));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) mx[q] = fmax3(mx[q], __uint_as_float(s[120 + 2 * q]), __uint_as_float(s[121 + 2 * q]));
            const float mx0 = fmax3(mx[0], mx[1], mx[2]), mx1 = fmax3(mx[3], mx[4], mx[5]), mx2 = fmaxf(mx[6], mx[7]), mx3 = mx2;
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
            TR(3);
            const float alpha = ex2_approx((m_run - m_new) * c);    // first tile: exp2(-inf) = 0
            const float mc = m_new * c;
            // p = exp2(s*c - m*c) -> bf16 pairs; the row sum is taken in fp32 before rounding
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), c, -mc));
                const float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), c, -mc));
                const float p2 = ex2_approx(fmaf(__uint_as_float(s[i + 2]), c, -mc));
                const float p3 = ex2_approx(fmaf(__uint_as_float(s[i + 3]), c, -mc));
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1);
                const __nv_bfloat162 h23 = __floats2bfloat162_rn(p2, p3);
                l0 += p0;
                l1 += p1;
                l2 += p2;
                l3 += p3;
                s[i >> 1] = *reinterpret_cast<const uint32_t*>(&h01);          // pack in place: s[0..63] = P
                s[(i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
            }
            TR(4);
            // P must have been consumed by the PV MMA of the previous tile: that is the same commit that publishes
            // O(it-1), so wait for it once here and fold the previous tile's O into the accumulator right away
            if (j > 0) accumulate_O(it - 1, alpha_prev);
            TR(5);
            // P -> TMEM: thread = row, column i = keys (2i, 2i+1) as a bf16 pair
            tmem_st_32x32(tP, &s[0]);
            tmem_st_32x32(tP + 32, &s[32]);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);
            TR(6);
            l_run = fmaf(l_run, alpha, (l0 + l1) + (l2 + l3));
            m_run = m_new;
            TR(7);
            alpha_prev = alpha;
        }
        accumulate_O(it - 1, alpha_prev);

        const int qpos = q0 + g * 128 + row;
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
        }   // work items
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

}  // namespace covo
