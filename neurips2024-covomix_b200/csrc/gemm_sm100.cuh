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
// main loop of tile i+1.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue
// (two warps per TMEM lane quarter, each owning half of the tile's columns).  Epilogue data path:
// TMEM -> registers (thread = row) -> [RoPE] -> + bias -> + residual (sub-tile TMA-loaded into smem)
// -> fp32 sub-tile into 128B-swizzled smem -> TMA store; and/or activation -> bf16/fp16 -> smem -> TMA
// store.  All global traffic of the epilogue is bulk-asynchronous; TMA clips rows/columns that fall
// outside the output tensor.  ConvTranspose1d in the classic polyphase form (row q, column n -> sample
// stride*q + n/C - pad) does not form a box and uses a per-element store path; the vocoder plan re-indexes
// the weights into an aligned polyphase form wherever k - 2 pad == stride (hifigan.cuh), which does.
//
// The same kernel serves every dense contraction on the hot path: the velocity net's Linear layers
// (taps = 1), its U-Net skip combiner (taps = 2 over two activation slots), HiFi-GAN's dilated
// Conv1d (taps = kernel size, tap_row = k*dilation - pad) and ConvTranspose1d (polyphase:
// N = stride*Cout, taps = ceil(K/stride) (+1 in the aligned form), tap_row = -j).
#pragma once
#include "ptx.cuh"

namespace covo {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;                 // 64 x 2 B = 128 B = one swizzle row
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_MAX_TAPS = 16;
constexpr int GEMM_STAGE_A_BYTES = GEMM_BM * GEMM_BK * 2;
constexpr int GEMM_EPI_BUF_BYTES = 4096;    // per epilogue warp: [32 rows x 128 B], 128B-swizzled

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_LRELU = 2 };

struct GemmArgs {
    CUtensorMap tmA;                 // (k, row, z) box (64, 128, 1), SWIZZLE_128B
    CUtensorMap tmB;                 // (k, n)      box (64, BN),     SWIZZLE_128B
    CUtensorMap tmOutF;              // fp32 output   (n, q, z) box (32, 32, 1), SWIZZLE_128B   (if has_out_f32 && !scatter)
    CUtensorMap tmRes;               // fp32 residual (n, q, z) box (32, 32, 1), SWIZZLE_128B   (if has_residual)
    CUtensorMap tmOutH;              // 16-bit output (n, q, z) box (64, 32, 1), SWIZZLE_128B   (if has_out_h && !scatter)
    int rows;                        // GEMM rows (q) per z
    int Z;                           // number of z entries (batch items); 1 for Linear layers
    int n_tiles;                     // N_pad / BN
    int n_valid;                     // columns >= n_valid are not stored
    int taps;
    int kc_per_tap;                  // Ktap / 64
    int tap_row[GEMM_MAX_TAPS];
    int tap_z[GEMM_MAX_TAPS];
    // ---- epilogue
    const float* bias;               // [N_pad] or null
    int has_residual;                // residual added through tmRes (may alias the fp32 output)
    int has_out_f32;
    int has_out_h;
    int act_h;                       // ACT_NONE | ACT_GELU | ACT_LRELU   (applied to the 16-bit output only)
    float slope;                     // LeakyReLU slope
    int h_is_fp16;                   // 16-bit output format: 0 = bf16, 1 = fp16
    const float2* rope;              // [seq][32] (cos, sin) or null
    int rope_seq;                    // position = q % rope_seq
    int rope_cols;                   // columns < rope_cols are rotated (q and k of to_qkv)
    // ---- scatter path (ConvTranspose1d): element offset = z*out_zs + q*out_rs + n + out_off,
    //      valid iff 0 <= up_s*q + n/phase_w - up_p < t_out
    int scatter;
    long long out_zs, out_rs, out_off;
    int up_s, up_p, phase_w, t_out;
    int scatter_c_valid;             // only columns with (n % phase_w) < scatter_c_valid are stored (0: all): the ConvTranspose1d in
                                     // front of the fused last vocoder stage writes just the channels that stage reads
    float* out_f32;
    void* out_h;
    int reverse;                     // != 0: the persistent schedule walks the tiles from the last to the first (serpentine order
                                     // along a chain of kernels: what the previous kernel wrote last is read first, out of L2)
};

// Debug timeline of gemm_run (CTA 0 only; null in production): [0] entry, [1] barriers initialised, [2] first stage landed,
// [3] last MMA issued, [4] accumulator of the last tile complete (epilogue warp 0), [5] its stores issued, [6] stores complete
// [7] cycles the MMA thread spent waiting for a free accumulator buffer (one 8-slot record per gemm_run call, first 200 calls)
__device__ long long* g_gemm_trace = nullptr;
__device__ int g_gemm_trace_n = 0;
#define GTR(slot) do { if (g_gemm_trace != nullptr && blockIdx.x == 0 && gtr_base >= 0) g_gemm_trace[gtr_base + (slot)] = clock64(); } while (0)

// CG = 2: the CTA is half of a cta_group::2 pair (M = 256 per MMA): it stages its own 128 rows of A and HALF of the weight
// tile per K chunk (32 KB instead of 48 KB at BN = 256), so the ring is deeper and shared-memory / L2 operand traffic per
// FLOP drops by a third.
template <int BN, int CG = 1>
struct GemmCfg {
    static constexpr int STAGE_B_BYTES = BN / CG * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = GEMM_STAGE_A_BYTES + STAGE_B_BYTES;
    // Two staging buffers per epilogue warp wherever a stage can be spared for them (everything but the 48 KB stages of the
    // single-CTA BN = 256 tile): the fp32 residual sub-tile of step k+1 is in flight while step k is added and stored.
    // (Measured: `out` 67.0 -> 63.7 us with 5 stages + 2 buffers on the cta_group::2 BN = 256 tile, but every other layer GEMM
    // loses 2-4 % to the shallower ring, so that tile keeps 6 stages and one buffer; the vocoder's narrow tiles gain 2 %.)
    static constexpr int EPI_BUFS = BN < 256 ? 2 : 1;
    static constexpr int STAGES = CG == 2 ? (BN == 256 ? 6 : 6) : ((BN == 256) ? 4 : (BN == 128 ? 5 : 6));
    static constexpr int EPI_STAGING_BYTES = GEMM_EPI_WARPS * EPI_BUFS * GEMM_EPI_BUF_BYTES;
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 512 /*barriers*/ + 1024 /*align*/;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b, int is_fp16) {
    if (is_fp16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Activation + 16-bit packing of one row's 64-column chunk.  The activation is dispatched ONCE, outside the unrolled loops, and the
// 16-bit format is a compile-time constant: with the dispatch inside the loop every pair paid two or three uniform branches,
// and with two or three warps per sub-partition nothing hides a branch's ~25 cycles -- the clock64 trace of the epilogue showed
// 1 800 cycles between the accumulator load and the staging store of a plain bf16 chunk (32 conversions), 4 100 with GELU, which
// made every K = 1024 GEMM and every narrow vocoder tile epilogue-bound (profiles/r02_gemm_epilogue_trace.txt).
template <bool FP16>
__device__ __forceinline__ uint32_t pack_h2c(float a, float b) {
    if (FP16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <bool FP16>
__device__ __forceinline__ void act_pack_chunk(uint32_t (&pk)[32], const uint32_t (&r)[64], int act_h, float slope) {
    if (act_h == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float o0, o1;
            gelu_fast2(o0, o1, __uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            pk[j] = pack_h2c<FP16>(o0, o1);
        }
    } else if (act_h == ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float a = __uint_as_float(r[2 * j]), b = __uint_as_float(r[2 * j + 1]);
            pk[j] = pack_h2c<FP16>(fmaxf(a, a * slope), fmaxf(b, b * slope));          // 0 < slope < 1
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = pack_h2c<FP16>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
    }
}

// The persistent tile loop of one GEMM launch, callable from the stand-alone kernel below and from the persistent
// flow-step kernel (flow_persistent.cuh), which walks a whole list of ops inside one launch.
//   desc : where the TMA descriptors live (kernel parameter space or global memory -- never shared memory)
//   args : the scalar fields (may be the same object, or a shared-memory copy of it)
//   smem : 1024-byte aligned, GemmCfg<BN>::SMEM_BYTES - 1024 bytes;  tmem_base: 2*BN allocated TMEM columns
//   epi_warp0 : first of the 8 epilogue warps (its TMEM lane quarter is warp % 4 whatever the offset)
//   cta / n_ctas : this CTA's slot in the persistent schedule
// Every thread of the CTA must call it (it initialises the mbarriers and synchronises the CTA once).
// ROLE: 0 = dispatch on the warp index; 1 = the caller is a control warp (TMA / MMA); 2 = the caller is an epilogue warp.
// MC: 1 = independent CTAs.  2 = the CTA is one of a cluster pair (launch with cluster dims (2,1,1)): the pair works on two
//     vertically adjacent M tiles of the same N tile and SHARES the weight tile -- each CTA fetches half of it and
//     multicasts that half into both CTAs' shared memory, so the W operand crosses the L2 -> SM fabric once per pair
//     (the K loop of a 128 x 256 tile asks the fabric for 96 B/clk/SM; chip-wide that is more than L2 delivers).
//     tmB's box is then (64, BN/2).  A stage may only be refilled when BOTH CTAs' MMAs have read it: every MMA commit on
//     empty[] is multicast to the pair (count 2).
// CG: 2 = cta_group::2 pair (launch with cluster dims (2,1,1), TMEM allocated with tmem_alloc_cg2 by both CTAs): same pair-tile
//     schedule as MC = 2, but ONE MMA of M = 256 per K16 step issued by the even-rank CTA; each CTA keeps only its own half
//     of the weight tile; tmB's box is (64, BN/2).  Requires MC = 1.
template <int BN, int AB_FMT /*0 fp16, 1 bf16*/, int ROLE = 0, int MC_ = 1, int CG = 1>
__device__ __forceinline__ void gemm_run(const GemmArgs* desc, const GemmArgs& args, uint8_t* smem, uint32_t tmem_base,
                                         int epi_warp0, int cta, int n_ctas) {
    static_assert(CG == 1 || MC_ == 1, "cta_group::2 and the multicast pair are alternatives");
    constexpr int MC = CG == 2 ? 2 : MC_;              // pair-tile schedule in both cases
    constexpr bool MCAST = MC_ == 2;                   // weight halves exchanged by TMA multicast (cta_group::1 MMAs)
    using Cfg = GemmCfg<BN, CG>;
    uint8_t* stage_base = smem;
    uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::EPI_STAGING_BYTES);
    uint64_t* full = bars;                         // [STAGES]
    uint64_t* empty = bars + Cfg::STAGES;          // [STAGES]
    uint64_t* tfull = bars + 2 * Cfg::STAGES;      // [2]
    uint64_t* tempty = bars + 2 * Cfg::STAGES + 2; // [2]
    uint64_t* rbar = bars + 2 * Cfg::STAGES + 4;   // [GEMM_EPI_WARPS * EPI_BUFS] residual sub-tile landed

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // BN >= 128: the 8 epilogue warps share every tile (two per TMEM lane quarter, half of the columns each).  BN = 64: a tile has
    // one 64-column chunk per lane quarter, so the warps form two SETS of four that take alternate tiles -- set s drains
    // accumulator buffer s -- and two tiles' epilogues run side by side (the narrow vocoder stages are epilogue-bound: a
    // 128 x 64 x 192 main loop is ~400 cycles against a ~4 000-cycle epilogue; with one set the 62-channel convs took 123-222 us).
    constexpr int EPI_WARPS_ACTIVE = (BN >= 128) ? GEMM_EPI_WARPS : 4;        // arrivals per accumulator buffer
    constexpr int EPI_TSTEP = (BN >= 128) ? 1 : 2;                            // tiles between two visits of an epilogue warp

    // MC == 2: the schedule runs over PAIR tiles (two M tiles x one N tile); `cta` / `n_ctas` then count clusters, and
    // m_tiles_per_z counts pairs (an odd tail pair has a phantom second tile: TMA zero-fills it, nothing is stored)
    const int crank = MC == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    if (MC == 2) {
        cta >>= 1;
        n_ctas >>= 1;
    }
    const int m_tiles_per_z = ((args.rows + GEMM_BM - 1) / GEMM_BM + MC - 1) / MC;
    const int total_tiles = args.Z * m_tiles_per_z * args.n_tiles;
    const int k_iters = args.taps * args.kc_per_tap;

    volatile int& s_gtr_base = *reinterpret_cast<volatile int*>(bars + 48);     // debug slot inside the 512-byte barrier region
    if (threadIdx.x == 0) {
        int b = -1;
        if (g_gemm_trace != nullptr && blockIdx.x == 0) {
            b = g_gemm_trace_n++;
            b = b < 200 ? 8 * b : -1;
        }
        s_gtr_base = b;
        if (b >= 0) g_gemm_trace[b] = clock64();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&desc->tmA);
        tma_prefetch_desc(&desc->tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MCAST ? 2 : 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], EPI_WARPS_ACTIVE * CG);      // CG = 2: the leader's barrier collects both CTAs' epilogue warps
        }
        for (int w = 0; w < GEMM_EPI_WARPS * Cfg::EPI_BUFS; ++w) mbar_init(&rbar[w], 1);
        fence_mbar_init();
    }
    if (MC == 2) cluster_sync_all();        // the peer's multicast may signal this CTA's barriers from now on
    else __syncthreads();
    const int gtr_base = s_gtr_base;
    if (threadIdx.x == 0) GTR(1);

    if (ROLE != 2 && warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = cta; tile < total_tiles; tile += n_ctas) {
                const int tt = args.reverse ? total_tiles - 1 - tile : tile;
                const int n_idx = tt % args.n_tiles;
                const int mz = tt / args.n_tiles;
                const int m_idx = (mz % m_tiles_per_z) * MC + crank;
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
                        if (CG == 2) {
                            // both CTAs' bytes are counted on the LEADER's barrier (its MMA thread is the only consumer)
                            if (crank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
                            const uint32_t lbar = mapa_u32(smem_u32(&full[s]), 0);
                            tma_load_3d_cg2(sa, &desc->tmA, lbar, kc * GEMM_BK, arow, az);
                            tma_load_2d_cg2(sb, &desc->tmB, lbar, (tap * args.kc_per_tap + kc) * GEMM_BK, n0 + crank * (BN / 2));
                            if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                            continue;
                        }
                        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                        tma_load_3d(sa, &desc->tmA, &full[s], kc * GEMM_BK, arow, az);
                        if (MCAST)          // this CTA's half of the weight tile, into both CTAs of the pair
                            tma_load_2d_mc(sb + crank * (Cfg::STAGE_B_BYTES / 2), &desc->tmB, &full[s],
                                           (tap * args.kc_per_tap + kc) * GEMM_BK, n0 + crank * (BN / 2), 0x3);
                        else
                            tma_load_2d(sb, &desc->tmB, &full[s], (tap * args.kc_per_tap + kc) * GEMM_BK, n0);
                        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (ROLE != 2 && warp == 1) {
        // ===================================================== MMA issuer (elect.sync: ptxas keeps operands in uniform registers)
        if ((CG == 1 || crank == 0) && elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM * CG, BN, AB_FMT, 0, 0);
            // descriptors of stage 0, built once; per stage / per K16 step only the start-address field is advanced
            const uint64_t dA0 = smem_desc_sw128(smem_u32(stage_base), 1024, 16);
            const uint64_t dB0 = smem_desc_sw128(smem_u32(stage_base) + GEMM_STAGE_A_BYTES, 1024, 16);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            const bool tracing = g_gemm_trace != nullptr && blockIdx.x == 0 && gtr_base >= 0;
            long long acc_stall = 0;                   // debug: cycles this thread waited for an accumulator buffer (epilogue-bound?)
            for (int tile = cta; tile < total_tiles; tile += n_ctas) {
                if (tracing) {
                    const long long t0 = clock64();
                    mbar_wait(&tempty[as], aph ^ 1);
                    acc_stall += clock64() - t0;
                    g_gemm_trace[gtr_base + 7] = acc_stall;
                } else {
                    mbar_wait(&tempty[as], aph ^ 1);
                }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(&full[s], ph);
                    if (it == 0 && tile == cta) GTR(2);
                    tc_fence_after();
                    const uint64_t da = desc_advance(dA0, s * Cfg::STAGE_BYTES);
                    const uint64_t db = desc_advance(dB0, s * Cfg::STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        if (CG == 2) umma_f16_cg2(d_tmem, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (it | k) != 0);
                        else umma_f16(d_tmem, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (it | k) != 0);
                    }
                    if (CG == 2) umma_commit_cg2_mc(&empty[s], 0x3);
                    else if (MCAST) umma_commit_mc(&empty[s], 0x3);
                    else umma_commit(&empty[s]);
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                }
                if (CG == 2) umma_commit_cg2_mc(&tfull[as], 0x3);
                else umma_commit(&tfull[as]);
                GTR(3);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else if (ROLE != 1 && warp >= epi_warp0 && warp - epi_warp0 < GEMM_EPI_WARPS) {
        // ===================================================== epilogue warps
        // warp e = warp - epi_warp0: TMEM lane quarter lq = warp % 4 (hardware restriction), column half = e / 4.
        const int e = warp - epi_warp0;
        const int lq = warp & 3;
        const int chalf = (BN >= 128) ? (e >> 2) : 0;
        const int eset = (BN >= 128) ? 0 : (e >> 2);          // BN = 64: which of the two warp sets (== accumulator buffer)
        constexpr int NB = Cfg::EPI_BUFS;
        uint8_t* buf = staging + e * NB * GEMM_EPI_BUF_BYTES;     // NB x [32 rows][128 B]; 16-B chunk c of row r at (c ^ (r&7))*16
        uint8_t* my_row = buf + lane * 128;
        const uint32_t my_row_s = smem_u32(my_row);
        const int sw = lane & 7;
        uint64_t* my_rbar = &rbar[e * NB];
        uint32_t rph = 0;                                         // NB = 2: bit j = parity of my_rbar[j]
        // NB = 2: fp32 sub-tile step k of this warp uses buffer k & 1; `pre`: the residual of step kstep has been requested
        uint32_t kstep = 0, kh = 0;
        bool pre = false;
        int as = eset;
        uint32_t aph = 0;
        constexpr int COLS_PER_WARP = (BN >= 128) ? BN / 2 : BN;
        constexpr int NCH = COLS_PER_WARP / 64;
        const bool use_tma = args.scatter == 0;
        const bool has_res = args.has_residual != 0;
        for (int tile = cta + eset * n_ctas; tile < total_tiles; tile += EPI_TSTEP * n_ctas) {
            const int tt = args.reverse ? total_tiles - 1 - tile : tile;
            const int n_idx = tt % args.n_tiles;
            const int mz = tt / args.n_tiles;
            const int m_idx = (mz % m_tiles_per_z) * MC + crank;
            const int z = mz / m_tiles_per_z;
            const int q_warp0 = m_idx * GEMM_BM + lq * 32;
            const int n0 = n_idx * BN + chalf * COLS_PER_WARP;
            const bool rows_live = q_warp0 < args.rows;            // warp-uniform: does this warp own any valid row?

            // residual sub-tile of the first fp32 sub-pass: fetched while the MMAs of this tile still run
            if (NB == 1) {
                if (has_res && rows_live && n0 < args.n_valid && lane == 0) {
                    bulk_wait_read0();
                    mbar_expect_tx(my_rbar, GEMM_EPI_BUF_BYTES);
                    tma_load_3d(buf, &desc->tmRes, my_rbar, n0, q_warp0, z);
                }
            } else if (has_res && use_tma && !pre && rows_live && n0 < args.n_valid) {
                if (lane == 0) {
                    const int j = kstep & 1;
                    bulk_wait_read0();
                    mbar_expect_tx(&my_rbar[j], GEMM_EPI_BUF_BYTES);
                    tma_load_3d(buf + j * GEMM_EPI_BUF_BYTES, &desc->tmRes, &my_rbar[j], n0, q_warp0, z);
                }
                pre = true;
            }

            // RoPE factors of this thread's row: the same 32 (cos, sin) pairs serve every head of the tile, and the row is known
            // before the accumulator is -- 16 independent 16-byte loads issued under the tail of the MMAs (per-chunk 8-byte
            // loads after the accumulator arrived cost 22k cycles per tile on the epilogue's critical path)
            float4 cs4[16];
            const bool use_rope = args.rope != nullptr && n0 < args.rope_cols && rows_live;
            if (use_rope) {
                const int pos = (q_warp0 + lane) % args.rope_seq;
                const float4* tab = reinterpret_cast<const float4*>(args.rope + static_cast<size_t>(pos) * 32);
#pragma unroll
                for (int j = 0; j < 16; ++j) cs4[j] = __ldg(tab + j);
            }
            // Bias of the first chunk, fetched into the same 64 registers while the MMAs of the tile still run (no op has both
            // RoPE and a bias); the next chunk's bias is requested as soon as this one has been added and flies under the
            // activation and the store.  Loaded after the accumulator arrived, the 16 broadcast loads cost ~800 cycles per chunk on
            // the epilogue's critical path (profiles/r02_gemm_epilogue_trace.txt) -- what kept FF1 epilogue-bound.
            const bool pre_bias = args.bias != nullptr && args.rope == nullptr && rows_live;
            if (pre_bias && n0 < args.n_valid) {
                const float4* b4 = reinterpret_cast<const float4*>(args.bias + n0);
#pragma unroll
                for (int j = 0; j < 16; ++j) cs4[j] = __ldg(b4 + j);
            }
            mbar_wait(&tfull[as], aph);
            if (e == 0 && lane == 0) GTR(4);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + static_cast<uint32_t>(as * BN + chalf * COLS_PER_WARP) +
                                   (static_cast<uint32_t>(lq * 32) << 16);

#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                uint32_t r[64];
                {
                    uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
                    uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[32]);
                    tmem_ld_32x32(t_acc + c * 64, r0);
                    tmem_ld_32x32(t_acc + c * 64 + 32, r1);
                    tmem_ld_wait();
                }
                if (c == NCH - 1) {
                    // accumulator fully read by this warp: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));
                        else mbar_arrive(&tempty[as]);
                    }
                }
                const int ncol0 = n0 + c * 64;
                if (!rows_live || ncol0 >= args.n_valid) continue;
                if (use_rope && ncol0 < args.rope_cols) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float4 cs = cs4[j >> 1];       // (cos_j, sin_j, cos_j+1, sin_j+1)
                        const float a1 = __uint_as_float(r[j]), a2 = __uint_as_float(r[j + 32]);
                        const float b1 = __uint_as_float(r[j + 1]), b2 = __uint_as_float(r[j + 33]);
                        r[j] = __float_as_uint(a1 * cs.x - a2 * cs.y);
                        r[j + 32] = __float_as_uint(a2 * cs.x + a1 * cs.y);
                        r[j + 1] = __float_as_uint(b1 * cs.z - b2 * cs.w);
                        r[j + 33] = __float_as_uint(b2 * cs.z + b1 * cs.w);
                    }
                }
                if (pre_bias) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 b = cs4[j];
                        float s0, s1, s2, s3;
                        add2(s0, s1, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), b.x, b.y);
                        add2(s2, s3, __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]), b.z, b.w);
                        r[4 * j] = __float_as_uint(s0);
                        r[4 * j + 1] = __float_as_uint(s1);
                        r[4 * j + 2] = __float_as_uint(s2);
                        r[4 * j + 3] = __float_as_uint(s3);
                    }
                    if (c + 1 < NCH && ncol0 + 64 < args.n_valid) {
                        const float4* b4 = reinterpret_cast<const float4*>(args.bias + ncol0 + 64);
#pragma unroll
                        for (int j = 0; j < 16; ++j) cs4[j] = __ldg(b4 + j);
                    }
                } else if (args.bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(args.bias + ncol0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 b = __ldg(b4 + j);
                        float s0, s1, s2, s3;
                        add2(s0, s1, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), b.x, b.y);
                        add2(s2, s3, __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]), b.z, b.w);
                        r[4 * j] = __float_as_uint(s0);
                        r[4 * j + 1] = __float_as_uint(s1);
                        r[4 * j + 2] = __float_as_uint(s2);
                        r[4 * j + 3] = __float_as_uint(s3);
                    }
                }
                if (use_tma && NB == 2) {
                    // ---------------- two staging buffers per warp.  fp32 sub-tile step k: buffer k & 1; at its start the residual
                    // of step k+1 (same chunk, next chunk or this warp's next tile) is requested into the other buffer, so a
                    // residual load is in flight while the previous sub-tile is added, written back and stored.
                    if (args.has_out_f32) {
#pragma unroll
                        for (int sp = 0; sp < 2; ++sp) {
                            const int ncol = ncol0 + sp * 32;
                            if (ncol < args.n_valid && (args.scatter_c_valid == 0 || ncol % args.phase_w < args.scatter_c_valid)) {
                                const int j = kstep & 1;
                                uint8_t* bj = buf + j * GEMM_EPI_BUF_BYTES;
                                const uint32_t row_s = my_row_s + j * GEMM_EPI_BUF_BYTES;
                                if (has_res) {
                                    int nq = q_warp0, nz = z, nn = -1;                 // the next fp32 sub-tile of this warp
                                    if (sp == 0 && ncol + 32 < args.n_valid) nn = ncol + 32;
                                    else if (c + 1 < NCH && ncol0 + 64 < args.n_valid) nn = ncol0 + 64;
                                    else if (tile + EPI_TSTEP * n_ctas < total_tiles) {
                                        const int t2 = args.reverse ? total_tiles - 1 - (tile + EPI_TSTEP * n_ctas) : tile + EPI_TSTEP * n_ctas;
                                        const int mz2 = t2 / args.n_tiles;
                                        nq = ((mz2 % m_tiles_per_z) * MC + crank) * GEMM_BM + lq * 32;
                                        nz = mz2 / m_tiles_per_z;
                                        const int n2 = (t2 % args.n_tiles) * BN + chalf * COLS_PER_WARP;
                                        if (nq < args.rows && n2 < args.n_valid) nn = n2;
                                    }
                                    if (lane == 0) {
                                        bulk_wait_read0();                              // both buffers' stores have drained
                                        if (!pre) {
                                            mbar_expect_tx(&my_rbar[j], GEMM_EPI_BUF_BYTES);
                                            tma_load_3d(bj, &desc->tmRes, &my_rbar[j], ncol, q_warp0, z);
                                        }
                                        if (nn >= 0) {
                                            mbar_expect_tx(&my_rbar[j ^ 1], GEMM_EPI_BUF_BYTES);
                                            tma_load_3d(buf + (j ^ 1) * GEMM_EPI_BUF_BYTES, &desc->tmRes, &my_rbar[j ^ 1], nn, nq, nz);
                                        }
                                    }
                                    pre = nn >= 0;
                                    mbar_wait(&my_rbar[j], (rph >> j) & 1u);
                                    rph ^= 1u << j;
#pragma unroll
                                    for (int c4 = 0; c4 < 8; ++c4) {
                                        const float4 v = lds128f(row_s + ((c4 ^ sw) << 4));
                                        const int b = sp * 32 + 4 * c4;
                                        float s0, s1, s2, s3;
                                        add2(s0, s1, __uint_as_float(r[b]), __uint_as_float(r[b + 1]), v.x, v.y);
                                        add2(s2, s3, __uint_as_float(r[b + 2]), __uint_as_float(r[b + 3]), v.z, v.w);
                                        r[b] = __float_as_uint(s0);
                                        r[b + 1] = __float_as_uint(s1);
                                        r[b + 2] = __float_as_uint(s2);
                                        r[b + 3] = __float_as_uint(s3);
                                    }
                                } else if (lane == 0) {
                                    bulk_wait_read1();          // the store that read this buffer (two stores ago) has drained
                                }
                                __syncwarp();
#pragma unroll
                                for (int c4 = 0; c4 < 8; ++c4) {
                                    const int b = sp * 32 + 4 * c4;
                                    sts128(row_s + ((c4 ^ sw) << 4), r[b], r[b + 1], r[b + 2], r[b + 3]);
                                }
                                fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_3d(&desc->tmOutF, bj, ncol, q_warp0, z);
                                    bulk_commit();
                                }
                                ++kstep;
                            }
                        }
                    }
                    if (args.has_out_h) {
                        // 16-bit [32 x 64] sub-tile: through the buffer the last fp32 step used (the other one may hold a
                        // prefetched residual); without fp32 steps the two buffers simply alternate
                        uint32_t pk[32];
                        act_pack_chunk<AB_FMT == 0>(pk, r, args.act_h, args.slope);
                        int j;
                        if (args.has_out_f32) {
                            j = (kstep - 1) & 1;
                            if (lane == 0) bulk_wait_read0();
                        } else {
                            j = kh++ & 1;
                            if (lane == 0) bulk_wait_read1();
                        }
                        const uint32_t row_s = my_row_s + j * GEMM_EPI_BUF_BYTES;
                        __syncwarp();
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4)
                            sts128(row_s + ((c4 ^ sw) << 4), pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&desc->tmOutH, buf + j * GEMM_EPI_BUF_BYTES, ncol0, q_warp0, z);
                            bulk_commit();
                        }
                    }
                } else if (use_tma) {
                    // ---------------- fp32 output (+ residual): two [32 x 32] fp32 sub-tiles
                    if (args.has_out_f32) {
#pragma unroll
                        for (int sp = 0; sp < 2; ++sp) {
                            const int ncol = ncol0 + sp * 32;
                            if (ncol < args.n_valid && (args.scatter_c_valid == 0 || ncol % args.phase_w < args.scatter_c_valid)) {
                                if (has_res) {
                                    mbar_wait(my_rbar, rph);
                                    rph ^= 1;
#pragma unroll
                                    for (int c4 = 0; c4 < 8; ++c4) {
                                        const float4 v = lds128f(my_row_s + ((c4 ^ sw) << 4));
                                        const int b = sp * 32 + 4 * c4;
                                        float s0, s1, s2, s3;
                                        add2(s0, s1, __uint_as_float(r[b]), __uint_as_float(r[b + 1]), v.x, v.y);
                                        add2(s2, s3, __uint_as_float(r[b + 2]), __uint_as_float(r[b + 3]), v.z, v.w);
                                        r[b] = __float_as_uint(s0);
                                        r[b + 1] = __float_as_uint(s1);
                                        r[b + 2] = __float_as_uint(s2);
                                        r[b + 3] = __float_as_uint(s3);
                                    }
                                } else if (lane == 0) {
                                    bulk_wait_read0();              // previous TMA store has finished reading the buffer
                                }
                                __syncwarp();
#pragma unroll
                                for (int c4 = 0; c4 < 8; ++c4) {
                                    const int b = sp * 32 + 4 * c4;
                                    sts128(my_row_s + ((c4 ^ sw) << 4), r[b], r[b + 1], r[b + 2], r[b + 3]);
                                }
                                fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_3d(&desc->tmOutF, buf, ncol, q_warp0, z);
                                    bulk_commit();
                                    if (has_res && sp == 0 && ncol + 32 < args.n_valid) {
                                        bulk_wait_read0();
                                        mbar_expect_tx(my_rbar, GEMM_EPI_BUF_BYTES);
                                        tma_load_3d(buf, &desc->tmRes, my_rbar, ncol + 32, q_warp0, z);
                                    }
                                }
                            }
                        }
                    }
                    // ---------------- 16-bit output: one [32 x 64] sub-tile
                    if (args.has_out_h) {
                        uint32_t pk[32];
                        act_pack_chunk<AB_FMT == 0>(pk, r, args.act_h, args.slope);
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4)
                            sts128(my_row_s + ((c4 ^ sw) << 4), pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&desc->tmOutH, buf, ncol0, q_warp0, z);
                            bulk_commit();
                        }
                    }
                    // residual sub-tile of the next chunk's first sub-pass
                    if (has_res && lane == 0 && c + 1 < NCH && ncol0 + 64 < args.n_valid) {
                        bulk_wait_read0();
                        mbar_expect_tx(my_rbar, GEMM_EPI_BUF_BYTES);
                        tma_load_3d(buf, &desc->tmRes, my_rbar, ncol0 + 64, q_warp0, z);
                    }
                } else {
                    // ---------------- scatter path (ConvTranspose1d): smem transpose, lane = column, per-element stores
                    const float* stg = reinterpret_cast<const float*>(buf);
#pragma unroll
                    for (int sp = 0; sp < 2; ++sp) {
                        __syncwarp();
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const int b = sp * 32 + 4 * c4;
                            sts128(my_row_s + ((c4 ^ sw) << 4), r[b], r[b + 1], r[b + 2], r[b + 3]);
                        }
                        __syncwarp();
                        const int n = ncol0 + sp * 32 + lane;
                        if (n < args.n_valid && (args.scatter_c_valid == 0 || n % args.phase_w < args.scatter_c_valid)) {
                            const int phase = n / args.phase_w;
                            const long long base = static_cast<long long>(z) * args.out_zs + args.out_off + n;
#pragma unroll 4
                            for (int rr = 0; rr < 32; ++rr) {
                                const int q = q_warp0 + rr;
                                const int t = args.up_s * q + phase - args.up_p;
                                if (t < 0 || t >= args.t_out) continue;
                                const float v = stg[rr * 32 + (((lane >> 2) ^ (rr & 7)) << 2) + (lane & 3)];
                                const long long off = base + static_cast<long long>(q) * args.out_rs;
                                if (args.out_f32 != nullptr) args.out_f32[off] = v;
                                if (args.out_h != nullptr) {
                                    float o = v;
                                    if (args.act_h == ACT_GELU) o = gelu_fast(o);
                                    else if (args.act_h == ACT_LRELU) o = lrelu(o, args.slope);
                                    static_cast<uint16_t*>(args.out_h)[off] =
                                        args.h_is_fp16 ? __half_as_ushort(__float2half_rn(o))
                                                       : __bfloat16_as_ushort(__float2bfloat16(o));
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (BN >= 128) {
                if (++as == 2) { as = 0; aph ^= 1; }
            } else {
                aph ^= 1;                                  // this set's buffer is used by every second tile
            }
        }
        if (e == 0 && lane == 0) GTR(5);
        if (lane == 0) bulk_wait_all();      // smem must stay valid (and the output be complete) when the caller moves on
        if (e == 0 && lane == 0) GTR(6);
    }
    if (MC == 2) cluster_sync_all();        // the peer may still be multicasting into this CTA's shared memory
}

template <int BN, int AB_FMT /*0 fp16, 1 bf16*/>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) {
        tmem_alloc(&tmem_slot, GemmCfg<BN>::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    gemm_run<BN, AB_FMT>(&args, args, smem, tmem_base, 2, blockIdx.x, gridDim.x);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, GemmCfg<BN>::TMEM_COLS);
    }
}

// Cluster-pair variant (see MC above): grid must be even, launched with the static cluster shape (2, 1, 1).
template <int BN, int AB_FMT /*0 fp16, 1 bf16*/>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1) gemm_tc_pair_kernel(const __grid_constant__ GemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) {
        tmem_alloc(&tmem_slot, GemmCfg<BN>::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    gemm_run<BN, AB_FMT, 0, 2>(&args, args, smem, tmem_base, 2, blockIdx.x, gridDim.x);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, GemmCfg<BN>::TMEM_COLS);
    }
}

// cta_group::2 variant (see CG above): grid must be even, static cluster shape (2, 1, 1); TMEM is allocated for the pair.
template <int BN, int AB_FMT /*0 fp16, 1 bf16*/>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1) gemm_tc_cg2_kernel(const __grid_constant__ GemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) {
        tmem_alloc_cg2(&tmem_slot, GemmCfg<BN, 2>::TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    gemm_run<BN, AB_FMT, 0, 1, 2>(&args, args, smem, tmem_base, 2, blockIdx.x, gridDim.x);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, GemmCfg<BN, 2>::TMEM_COLS);
    }
}

}  // namespace covo
