// Non-causal, unmasked multi-head attention (dim_head = 64) for sm_100a:  O = softmax(Q K^T / 8) V.
//
// Replaces Attend.forward's einsum/softmax/einsum (covomix/covomix_model/attend.py:110-124), which
// materialises the [B,H,N,N] fp32 score tensor in HBM; here scores never leave the SM.
//
// Persistent: grid = min(#work items, #SMs).  A sequence is cut into 128-query groups; a work item is a PAIR of groups
// A and B of one (sequence, head) that share every K/V tile (or the single left-over group when the group count is odd;
// pairs are dealt first so the short items fill the tail of the schedule).  The TMA / MMA / softmax pipelines run
// straight across item boundaries (Q is double-buffered).  Q/K/V tiles are TMA-loaded straight out of the to_qkv
// GEMM's [B*N, 3*H*64] bf16 output (3-D tensor map: column, position, sequence; rows past the end of a sequence are
// zero-filled by TMA and masked to -inf in the softmax).
//   warp 0      : TMA producer (Q tiles of the next item, a 3-stage ring of K tiles and a 4-stage ring of V tiles)
//   warp 1      : tcgen05.mma issuer:  S_g = Q_g K_j^T (M128 x N128 x K64, K-major operands, into TMEM),
//                 O_g += P_g V_j (M128 x N64 x K128, A = P from TMEM, B = V MN-major); O ACCUMULATES IN TMEM over the
//                 key tiles of an item.  Issue order per step u:  S_A(u), PV_B(u-2), S_B(u), PV_A(u-1) -- the order in
//                 which the operands become ready when group B runs half a tile behind group A, so that one group's
//                 exponentials (MUFU-bound) overlap the other group's TMEM traffic, row max and barrier latencies.
//   warps 4-7   : softmax group A, warps 8-11: group B; one query row per thread (TMEM lane == row, so row max / row
//                 sum need no shuffles): S_j from TMEM into registers (buffer released at once) -> running max ->
//                 p = exp2((s - m) / 8 * log2 e) -> P_j as bf16 pairs into TMEM (tcgen05.st, the A operand of the PV MMA).
//                 The running max is only raised (and O, l rescaled in TMEM by the same thread) when it moves by more than
//                 2^8 for some row of the warp -- the result is the exact softmax either way, P merely carries a
//                 common factor <= 256 -- so the common per-tile path touches neither O nor a correction factor.
//                 Arithmetic: packed fp32 pairs (FFMA2 / FADD2); a fixed fraction of the exponentials is evaluated on
//                 the FMA pipe (Cody-Waite split + degree-3 minimax, |rel err| < 7.5e-5, far below the bf16 rounding
//                 of P) because MUFU.EX2 (8 cycles per warp instruction per sub-partition) is the binding unit:
//                 tools/micro/softmax_loop.cu, profiles/r02_softmax_loop_micro.txt.
// Registers are rebalanced with setmaxnreg (producer/MMA warpgroup 56, softmax warpgroups 224; the pool of setmaxnreg.inc is exactly what warpgroup 0 releases)..
#pragma once
#include <type_traits>
#include "ptx.cuh"

namespace covo {

// Optional clock64 timeline of the softmax loop (tools/micro/attn_trace.cu); compiled out in the library.
#ifdef ATT_TRACE
__device__ long long g_trace[4096];
#define TR(slot) do { if (blockIdx.x == 0 && lane == 0 && itg >= 40 && itg < 44) g_trace[(warp) * 256 + (itg - 40) * 16 + (slot)] = clock64(); } while (0)
#else
#define TR(slot)
#endif

constexpr int ATT_BM = 256;      // queries per pair item (two groups of 128)
constexpr int ATT_BG = 128;      // queries per group
constexpr int ATT_BN = 128;      // keys per iteration
constexpr int ATT_D = 64;        // dim_head
constexpr int ATT_KS = 3;        // K ring stages
constexpr int ATT_VS = 4;        // V ring stages (V_j is still needed two steps after K_j)
constexpr int ATT_THREADS = 384; // warpgroup 0: TMA + MMA (+2 idle warps); warpgroups 1, 2: softmax A, B
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;                 // 16 KB: one [128 x 64] bf16 tile
constexpr int ATT_SMEM_BYTES = (4 + ATT_KS + ATT_VS) * ATT_TILE_BYTES + 512 + 1024;
constexpr int ATT_TMEM_COLS = 512;                           // S: 2 x 128, O: 2 x 64, P (bf16 pairs): 2 x 64
constexpr float ATT_RESCALE_LOG2 = 8.0f;                     // raise the running max only when it moves by more than 2^8

struct AttnArgs {
    CUtensorMap tmQKV;     // (col, pos, seq) over [Bt, N, 3*inner] bf16, box (64, 128, 1), SWIZZLE_128B
    __nv_bfloat16* out;    // [Bt, N, inner]
    int N;                 // sequence length
    int heads;
    int inner;             // heads * 64
    int n_pairs;           // pair items per (sequence, head) = ceil(N / 128) / 2
    int n_pair_items;      // n_pairs * heads * sequences
    int n_items;           // + one single-group item per (sequence, head) when ceil(N / 128) is odd
    int stagger;           // != 0: exponential token between the two query groups (anti-phase); 0: free running
    int sequences;         // Bt
    int reverse;           // != 0: sequences are walked from the last to the first (serpentine order along the kernel chain)
};

inline void attn_fill_items(AttnArgs& a, int sequences) {
    const int groups = (a.N + ATT_BG - 1) / ATT_BG;
    a.n_pairs = groups / 2;
    a.n_pair_items = a.n_pairs * a.heads * sequences;
    a.n_items = a.n_pair_items + (groups & 1) * a.heads * sequences;
    a.sequences = sequences;
}

struct AttnItem {
    int seq, head, q0, two;
};
__device__ __forceinline__ AttnItem attn_decode(const AttnArgs& args, int w) {
    AttnItem it;
    int r;
    if (w < args.n_pair_items) {
        it.q0 = (w % args.n_pairs) * ATT_BM;
        r = w / args.n_pairs;
        it.two = 1;
    } else {
        it.q0 = args.n_pairs * ATT_BM;
        r = w - args.n_pair_items;
        it.two = 0;
    }
    it.head = r % args.heads;
    it.seq = args.reverse ? args.sequences - 1 - r / args.heads : r / args.heads;
    return it;
}

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
// 2^x for a pair on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x) through the 1.5 * 2^23 trick, 2^f from a
// degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5), 2^n added into the exponent field.
// x is clamped at -125 (the result is then ~2e-38 instead of 0 -- masked keys carry -inf).  x <= ~8 by construction.
__device__ __forceinline__ void exp2_fma2(float& p0, float& p1, float x0, float x1) {
    constexpr float MAGIC = 12582912.f;
    x0 = fmaxf(x0, -125.f);
    x1 = fmaxf(x1, -125.f);
    float t0, t1, n0, n1, f0, f1, q0, q1;
    add2(t0, t1, x0, x1, MAGIC, MAGIC);
    add2(n0, n1, t0, t1, -MAGIC, -MAGIC);
    fma2(f0, f1, n0, n1, -1.f, -1.f, x0, x1);
    fma2(q0, q1, f0, f1, 0.055171459913f, 0.055171459913f, 0.24261085689f, 0.24261085689f);
    fma2(q0, q1, q0, q1, f0, f1, 0.69326096773f, 0.69326096773f);
    fma2(q0, q1, q0, q1, f0, f1, 0.9999281168f, 0.9999281168f);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// POLY_MASK: bit (pair index mod 8) set -> that pair of scores takes the FMA-pipe exponential (0: all on the MUFU;
// 0x88: one pair in four; 0x92: three in eight).
// Exponential token between the two query groups, one per SM sub-partition (named barriers 1-4: group A's warp arrives,
// group B's warp of the same sub-partition waits; 5-8 the other way round; 64 participants each): a warp starts the
// exponentials of a tile only when the other group's warp on its sub-partition (the one it shares the MUFU with) has
// passed score HANDOFF of its own tile.  This keeps the groups in ANTI-PHASE -- one group's TMEM round trips, row max
// and barrier latencies run under the other group's MUFU-bound loop.  Left alone the groups lock IN phase (both in the
// exponentials, then both outside them; profiles/r02_attention_trace.txt).
__device__ __forceinline__ void named_bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// The persistent work-item loop, callable from the stand-alone kernel below and from the persistent flow-step kernel
// (flow_persistent.cuh).  desc: where the TMA descriptor lives (parameter space or global memory); args: the scalar
// fields (may be a shared-memory copy); smem: 1024-byte aligned, ATT_SMEM_BYTES - 1024 bytes; tmem_base: 512 allocated
// columns.  Warp roles: 0 TMA, 1 MMA, 4-7 / 8-11 softmax groups A / B.  Every thread of the CTA must call it.
// ROLE: 0 = dispatch on the warp index (stand-alone kernel), 1 = caller is a control warp (0-3), 2 = caller is a softmax warp
// (4-11) -- the persistent kernel splits its warps once at the top so that ptxas sees each role's register budget.
// SETREG: apply setmaxnreg here (warps 0-3 -> 56, warps 4-11 -> 224); otherwise the caller has done it.
template <int POLY_MASK, int HANDOFF = 128, int ROLE = 0, bool SETREG = true>
__device__ __forceinline__ void attention_run(const AttnArgs* desc, const AttnArgs& args, uint8_t* smem, uint32_t tmem_base, int cta,
                                              int n_ctas) {
    uint8_t* sQ = smem;                                 // 2 buffers x 2 groups
    uint8_t* sK = sQ + 4 * ATT_TILE_BYTES;              // ATT_KS stages
    uint8_t* sV = sK + ATT_KS * ATT_TILE_BYTES;         // ATT_VS stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_VS * ATT_TILE_BYTES);
    uint64_t* q_full = bars + 0;                        // [2 buffers]
    uint64_t* q_empty = bars + 2;                       // [2 buffers]
    uint64_t* k_full = bars + 4;                        // [ATT_KS]
    uint64_t* k_empty = k_full + ATT_KS;
    uint64_t* v_full = k_empty + ATT_KS;                // [ATT_VS]
    uint64_t* v_empty = v_full + ATT_VS;
    uint64_t* s_full = v_empty + ATT_VS;                // [2 groups]  MMA -> softmax : S_g(t) is in TMEM
    uint64_t* s_free = s_full + 2;                      //             softmax -> MMA : S_g(t) has been pulled into registers
    uint64_t* p_full = s_free + 2;                      //             softmax -> MMA : P_g(t) is in TMEM (and O_g rescaled if needed)
    uint64_t* o_full = p_full + 2;                      //             MMA -> softmax : P_g(t) V(t) done (P consumed, O_g updated)
    uint64_t* mma_done = o_full + 2;                    // every MMA and commit of this call has completed (barriers may be re-initialised)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kv = (args.N + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&desc->tmQKV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
        }
        for (int i = 0; i < ATT_KS; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
        }
        for (int i = 0; i < ATT_VS; ++i) {
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        mbar_init(mma_done, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t tmem_S = tmem_base;            // + group * 128
    const uint32_t tmem_O = tmem_base + 256;      // + group * 64
    const uint32_t tmem_P = tmem_base + 384;      // + group * 64: P as bf16 pairs, the A operand of the PV MMA

    if (ROLE != 2 && warp < 4) {
        if (SETREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0 && elect_one()) {
            // ===================================================== TMA producer
            int ks = 0, vs = 0;
            uint32_t kph = 0, vph = 0;
            int il = 0;                                        // local work-item counter
            for (int w = cta; w < args.n_items; w += n_ctas, ++il) {
                const AttnItem item = attn_decode(args, w);
                const int qb = il & 1;
                mbar_wait(&q_empty[qb], ((il >> 1) & 1) ^ 1);
                mbar_expect_tx(&q_full[qb], (item.two ? 2 : 1) * ATT_TILE_BYTES);
                tma_load_3d(sQ + (2 * qb) * ATT_TILE_BYTES, &desc->tmQKV, &q_full[qb], item.head * ATT_D, item.q0, item.seq);
                if (item.two)
                    tma_load_3d(sQ + (2 * qb + 1) * ATT_TILE_BYTES, &desc->tmQKV, &q_full[qb], item.head * ATT_D, item.q0 + ATT_BG,
                                item.seq);
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&k_empty[ks], kph ^ 1);
                    mbar_expect_tx(&k_full[ks], ATT_TILE_BYTES);
                    tma_load_3d(sK + ks * ATT_TILE_BYTES, &desc->tmQKV, &k_full[ks], args.inner + item.head * ATT_D, j * ATT_BN,
                                item.seq);
                    if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
                    mbar_wait(&v_empty[vs], vph ^ 1);
                    mbar_expect_tx(&v_full[vs], ATT_TILE_BYTES);
                    tma_load_3d(sV + vs * ATT_TILE_BYTES, &desc->tmQKV, &v_full[vs], 2 * args.inner + item.head * ATT_D,
                                j * ATT_BN, item.seq);
                    if (++vs == ATT_VS) { vs = 0; vph ^= 1; }
                }
            }
        } else if (warp == 1 && elect_one()) {
            // ===================================================== MMA issuer (elect.sync: ptxas keeps operands in uniform registers)
            constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_BN, 1, 0, 0);   // Q K^T : both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 1, 0, 1);    // P V   : V is MN-major
            // Descriptors of buffer 0; per MMA only the 14-bit start-address field is advanced.
            const uint64_t dQ0 = smem_desc_sw128(smem_u32(sQ), 1024, 16);
            const uint64_t dK0 = smem_desc_sw128(smem_u32(sK), 1024, 16);
            const uint64_t dV0 = smem_desc_sw128(smem_u32(sV), 1024, ATT_TILE_BYTES);
            int n_my = 0;
            for (int w = cta; w < args.n_items; w += n_ctas) ++n_my;
            const uint32_t total = static_cast<uint32_t>(n_my) * n_kv;
            // cursor of step u (the S tile) and the descriptors of the two previous steps (their PV products)
            int il = 0, j = 0, w = cta;
            int two_c = n_my > 0 ? attn_decode(args, w).two : 0;
            int ks = 0;
            uint32_t kph = 0;
            int j1 = 0, two1 = 0, j2 = 0, two2 = 0;           // (j, two) of steps u-1 and u-2
            uint32_t cS0 = 0, cS1 = 0, cP0 = 0, cP1 = 0;      // S / PV products issued per group (barrier parities)
            for (uint32_t u = 0; u <= total + 1; ++u) {
                const bool have = u < total;
                const int qb = il & 1;
                // ---- S_A(u)
                if (have) {
                    if (j == 0) mbar_wait(&q_full[qb], (il >> 1) & 1);
                    mbar_wait(&k_full[ks], kph);
                    mbar_wait(&s_free[0], (cS0 & 1) ^ 1);          // softmax A has pulled S_A(u-1) into registers
                    tc_fence_after();
                    const uint64_t bQ = desc_advance(dQ0, (2 * qb) * ATT_TILE_BYTES);
                    const uint64_t bK = desc_advance(dK0, ks * ATT_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < ATT_D / 16; ++k)
                        umma_f16(tmem_S, desc_advance(bQ, k * 32), desc_advance(bK, k * 32), idesc_s, k != 0);
                    umma_commit(&s_full[0]);
                    ++cS0;
                }
                // ---- O_B += P_B(u-2) V(u-2)
                if (u >= 2 && u - 2 < total && two2) {
                    const uint32_t tt = u - 2;
                    const int vs = tt % ATT_VS;
                    mbar_wait(&v_full[vs], (tt / ATT_VS) & 1);
                    mbar_wait(&p_full[1], cP1 & 1);
                    tc_fence_after();
                    const uint64_t bV = desc_advance(dV0, vs * ATT_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < ATT_BN / 16; ++k)
                        umma_f16_ts(tmem_O + ATT_D, tmem_P + 64 + k * 8, desc_advance(bV, k * 2048), idesc_o, (k != 0) || (j2 != 0));
                    umma_commit(&o_full[1]);
                    umma_commit(&v_empty[vs]);
                    ++cP1;
                }
                // ---- S_B(u), then K(u) (and after the item's last tile its Q buffer) can be refilled
                if (have) {
                    if (two_c) {
                        mbar_wait(&s_free[1], (cS1 & 1) ^ 1);
                        tc_fence_after();
                        const uint64_t bQ = desc_advance(dQ0, (2 * qb + 1) * ATT_TILE_BYTES);
                        const uint64_t bK = desc_advance(dK0, ks * ATT_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < ATT_D / 16; ++k)
                            umma_f16(tmem_S + ATT_BN, desc_advance(bQ, k * 32), desc_advance(bK, k * 32), idesc_s, k != 0);
                        umma_commit(&s_full[1]);
                        ++cS1;
                    }
                    umma_commit(&k_empty[ks]);
                    if (j == n_kv - 1) umma_commit(&q_empty[qb]);
                    if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
                }
                // ---- O_A += P_A(u-1) V(u-1)
                if (u >= 1 && u - 1 < total) {
                    const uint32_t tt = u - 1;
                    const int vs = tt % ATT_VS;
                    mbar_wait(&v_full[vs], (tt / ATT_VS) & 1);
                    mbar_wait(&p_full[0], cP0 & 1);
                    tc_fence_after();
                    const uint64_t bV = desc_advance(dV0, vs * ATT_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < ATT_BN / 16; ++k)
                        umma_f16_ts(tmem_O, tmem_P + k * 8, desc_advance(bV, k * 2048), idesc_o, (k != 0) || (j1 != 0));
                    umma_commit(&o_full[0]);
                    if (!two1) umma_commit(&v_empty[vs]);          // single-group item: nobody else reads V(u-1)
                    ++cP0;
                }
                // ---- advance: step u becomes u-1, u-1 becomes u-2
                j2 = j1;
                two2 = two1;
                j1 = j;
                two1 = two_c;
                if (have && ++j == n_kv) {
                    j = 0;
                    ++il;
                    w += n_ctas;
                    if (w < args.n_items) two_c = attn_decode(args, w).two;
                }
            }
            umma_commit(mma_done);
            mbar_wait(mma_done, 0);
        }
    } else if (ROLE != 1 && warp >= 4 && warp < 12) {
        if (SETREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===================================================== softmax warpgroups: thread == query row
        const int g = (warp - 4) >> 2;                 // 0: group A (warps 4-7), 1: group B (warps 8-11)
        const int lq = warp & 3;                       // TMEM lane quarter == SM sub-partition of this warp
        const int row = lq * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lq * 32) << 16;
        const uint32_t tS = tmem_S + g * ATT_BN + lane_addr;
        const uint32_t tO = tmem_O + g * ATT_D + lane_addr;
        const uint32_t tP = tmem_P + g * 64 + lane_addr;
        constexpr float c = 0.125f * 1.4426950408889634f;   // dim_head^-0.5 * log2(e); dim_head is 64 in this library
        uint32_t itg = 0;            // key tiles this group has processed (barrier parities)
        uint32_t itp = 0;            // ... of which in pair items (token hand-offs with the other group)
        bool s_ready = false;        // s_full of the coming tile has already been observed complete

        for (int w = cta; w < args.n_items; w += n_ctas) {
            const AttnItem item = attn_decode(args, w);
            if (g == 1 && !item.two) continue;
            const bool ho = HANDOFF > 0 && args.stagger != 0 && item.two;
            float m_run = -INFINITY;     // reference maximum of the raw scores (lags the true running max by < 2^8 / c)
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;     // running sum of p (four partial sums)
            // One key tile.  MASKED: the sequence's last, partial tile (keys past N are -inf, every exponential on the MUFU) -- a
            // compile-time variant, so that the full tiles carry neither the test nor the second copy of the exponential loop
            // behind a run-time branch (uniform branches are not free here: one softmax warp per sub-partition and group).
            auto key_tile = [&](auto masked_c, const int j) {
                constexpr bool MASKED = decltype(masked_c)::value;
                TR(0);
                if (!s_ready) mbar_wait(&s_full[g], itg & 1);
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

                if (MASKED) {
                    const int kv_valid = args.N - j * ATT_BN;  // keys >= kv_valid are padding
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
                const float m_tile = fmaxf(fmaxf(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5])), fmaxf(mx[6], mx[7]));
                TR(3);
                // Raise the reference maximum only when some row of this warp would otherwise exceed 2^8 (warp-uniform
                // decision: the TMEM accesses of the rescale are warp-collective).  alpha = 1 for the rows that did not move.
                float alpha = 1.0f;
                bool rescale = false;
                if (j == 0) {
                    m_run = m_tile;                                 // O and l start from zero: nothing to rescale
                } else {
                    const float m_new = fmaxf(m_run, m_tile);
                    if (__any_sync(0xffffffffu, (m_new - m_run) * c > ATT_RESCALE_LOG2)) {
                        alpha = ex2_approx((m_run - m_new) * c);
                        m_run = m_new;
                        rescale = true;
                        mul2(l0, l1, l0, l1, alpha, alpha);
                        mul2(l2, l3, l2, l3, alpha, alpha);
                    }
                }
                const float mc = -m_run * c;
                // The P buffer is free (and O_g up to date) once the PV product of this group's previous tile has completed:
                // poll that barrier now, consume the answer half-way through the exponentials (even a completed mbarrier costs ~100 cycles)
                const bool o_ready = itg > 0 ? mbar_try_wait(&o_full[g], (itg - 1) & 1) : true;
                // token: A's pair tile p starts after B's hand-off point of tile p-1, B's tile p after A's of tile p
                if (ho && (g == 1 || itp > 0)) named_bar_sync64(g == 0 ? 5 + lq : 1 + lq);
                // p = exp2(s*c - m*c) -> bf16 pairs packed in place (s[0..63] = P); the row sum is taken in fp32 before rounding
                constexpr bool all_mufu = (POLY_MASK == 0) || MASKED;    // -inf scores only go through the MUFU
                if (all_mufu) {
#pragma unroll
                    for (int i = 0; i < 128; i += 4) {
                        if (i == HANDOFF && ho) named_bar_arrive64(g == 0 ? 1 + lq : 5 + lq);
                        if (i == 64) {                       // first half of P goes to TMEM under the second half's exponentials
                            if (!o_ready) mbar_wait(&o_full[g], (itg - 1) & 1);
                            tc_fence_after();
                            if (!rescale) tmem_st_32x32(tP, &s[0]);
                        }
                        if (i == 96) s_ready = (j + 1 < n_kv) ? mbar_try_wait(&s_full[g], (itg + 1) & 1) : false;
                        float x0, x1, x2, x3;
                        fma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), c, c, mc, mc);
                        fma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), c, c, mc, mc);
                        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
                        add2(l0, l1, l0, l1, p0, p1);
                        add2(l2, l3, l2, l3, p2, p3);
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1);
                        const __nv_bfloat162 h23 = __floats2bfloat162_rn(p2, p3);
                        s[i >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
                        s[(i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 128; i += 4) {
                        if (i == HANDOFF && ho) named_bar_arrive64(g == 0 ? 1 + lq : 5 + lq);
                        if (i == 64) {                       // first half of P goes to TMEM under the second half's exponentials
                            if (!o_ready) mbar_wait(&o_full[g], (itg - 1) & 1);
                            tc_fence_after();
                            if (!rescale) tmem_st_32x32(tP, &s[0]);
                        }
                        if (i == 96) s_ready = (j + 1 < n_kv) ? mbar_try_wait(&s_full[g], (itg + 1) & 1) : false;
                        float x0, x1, x2, x3, p0, p1, p2, p3;
                        fma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), c, c, mc, mc);
                        fma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), c, c, mc, mc);
                        if ((POLY_MASK >> ((i >> 1) & 7)) & 1) exp2_fma2(p0, p1, x0, x1);
                        else { p0 = ex2_approx(x0); p1 = ex2_approx(x1); }
                        if ((POLY_MASK >> (((i >> 1) + 1) & 7)) & 1) exp2_fma2(p2, p3, x2, x3);
                        else { p2 = ex2_approx(x2); p3 = ex2_approx(x3); }
                        add2(l0, l1, l0, l1, p0, p1);
                        add2(l2, l3, l2, l3, p2, p3);
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1);
                        const __nv_bfloat162 h23 = __floats2bfloat162_rn(p2, p3);
                        s[i >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
                        s[(i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    }
                }
                if (HANDOFF >= 128 && ho) named_bar_arrive64(g == 0 ? 1 + lq : 5 + lq);
                itp += item.two;
                TR(4);
                TR(5);
                if (rescale) {
                    uint32_t o[64];
                    uint32_t (&o0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[0]);
                    uint32_t (&o1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[32]);
                    tmem_ld_32x32(tO, o0);
                    tmem_ld_32x32(tO + 32, o1);
                    tmem_ld_wait();
#pragma unroll
                    for (int d = 0; d < ATT_D; d += 2) {
                        float a0, a1;
                        mul2(a0, a1, __uint_as_float(o[d]), __uint_as_float(o[d + 1]), alpha, alpha);
                        o[d] = __float_as_uint(a0);
                        o[d + 1] = __float_as_uint(a1);
                    }
                    tmem_st_32x32(tO, &o[0]);
                    tmem_st_32x32(tO + 32, &o[32]);
                    tmem_st_32x32(tP, &s[0]);
                }
                // P -> TMEM: thread = row, column i = keys (2i, 2i+1) as a bf16 pair (the first half is already on its way)
                tmem_st_32x32(tP + 32, &s[32]);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g]);
                TR(6);
            };
            const int n_full = (args.N % ATT_BN) ? n_kv - 1 : n_kv;      // tiles without padding keys
            for (int j = 0; j < n_full; ++j, ++itg) key_tile(std::false_type(), j);
            if (n_full < n_kv) {
                key_tile(std::true_type(), n_full);
                ++itg;
            }
            // ---- item epilogue: O_g / l from TMEM (the next item's first PV product, which overwrites O_g, is only issued
            // after this thread's next p_full arrival)
            mbar_wait(&o_full[g], (itg - 1) & 1);
            tc_fence_after();
            uint32_t o[64];
            {
                uint32_t (&o0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[0]);
                uint32_t (&o1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[32]);
                tmem_ld_32x32(tO, o0);
                tmem_ld_32x32(tO + 32, o1);
                tmem_ld_wait();
            }
            const int qpos = item.q0 + g * ATT_BG + row;
            if (qpos < args.N) {
                const float inv = 1.0f / ((l0 + l1) + (l2 + l3));
                __nv_bfloat16* dst = args.out + (static_cast<size_t>(item.seq) * args.N + qpos) * args.inner + item.head * ATT_D;
#pragma unroll
                for (int d = 0; d < ATT_D; d += 8) {
                    uint4 v;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(o[d]) * inv, __uint_as_float(o[d + 1]) * inv);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(o[d + 2]) * inv, __uint_as_float(o[d + 3]) * inv);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(o[d + 4]) * inv, __uint_as_float(o[d + 5]) * inv);
                    __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(o[d + 6]) * inv, __uint_as_float(o[d + 7]) * inv);
                    v.x = *reinterpret_cast<uint32_t*>(&h0);
                    v.y = *reinterpret_cast<uint32_t*>(&h1);
                    v.z = *reinterpret_cast<uint32_t*>(&h2);
                    v.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(dst + d) = v;
                }
            }
        }   // work items
    }

}

template <int POLY_MASK, int HANDOFF = 128>
__global__ void __launch_bounds__(ATT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) {
        tmem_alloc(&tmem_slot, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    attention_run<POLY_MASK, HANDOFF, 0, true>(&args, args, smem, tmem_base, blockIdx.x, gridDim.x);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

}  // namespace covo
