// Persistent flow-step kernel: the whole ODE loop of ConditionalFlowMatcherWrapper.sample (acoustic.py:623-657: every
// evaluation of the velocity net + CFG combine + Euler / midpoint update) in ONE cooperative launch.
//
// One CTA per SM stays resident and walks the op list of a network evaluation -- to_embed, conv-pos, and per layer
// [skip combiner,] AdaRMSNorm, to_qkv (+RoPE), attention, to_out (+residual), AdaRMSNorm, FF1 (+GELU), FF2 (+residual),
// then the final RMSNorm, to_pred and the CFG + solver update -- n_evals times, separated by grid barriers instead of
// kernel boundaries.  TMEM is allocated once, the warp roles never change (warp 0 TMA producer, warp 1 tcgen05.mma
// issuer, warps 4-11 epilogue / softmax / element-wise workers), and the ops reuse the tile loops of the stand-alone
// kernels (gemm_run, attention_run), so numerics are identical to the launch-per-op path.
//
// Why: for short utterances (BASELINE configs[1]: M = 2 x 650 rows) a GEMM is 44-352 tiles of a few microseconds each
// and the launch-per-op graph spends more time in launch gaps, TMEM allocation, barrier set-up and pipeline drains
// (66 launches per evaluation, ~13 us each) than in the tensor pipe.  Here an op boundary costs one grid barrier
// (~1.5 us).  Large batches (C3) keep the launch-per-op path: their kernels run for 100+ us each.
//
// Memory model across CTAs: a producer's global writes are either generic stores (element-wise ops) or bulk tensor
// stores whose completion the issuing lane waits for (cp.async.bulk.wait_group 0) before the CTA arrives at the barrier
// (red.release.gpu by one thread after __syncthreads); consumers read through TMA (L2) or with ld.global.cg -- never
// through L1, which is not coherent across SMs -- after an acquire on the barrier word and a proxy fence.
#pragma once
#include "common.cuh"

namespace covo {

enum : int { MOP_GEMM = 0, MOP_ATTN = 1, MOP_NORM = 2, MOP_CONVPOS = 3, MOP_CFG = 4 };

struct NormOpArgs {
    const float* x;
    const float* gamma;          // + eval * gb_stride
    const float* beta;           // may be null (final RMSNorm); + eval * gb_stride
    __nv_bfloat16* out;
    int M;
    long long gb_stride;         // floats between the AdaLN rows of consecutive evaluation times (0: time-independent)
};
struct ConvposOpArgs {
    const float* h;
    const float* wT;
    const float* bias;
    float* x;
    __nv_bfloat16* x_h;
    int N, D, Bt;
};
struct CfgOpArgs {
    const float* vpred;
    float* x_state;
    __nv_bfloat16* xin;
    int BN, dx, ldx;
    float s;
    int two_branch;
};
struct alignas(128) MegaOp {
    int type;
    int bn;                      // GEMM tile width
    int pad_[2];
    union {
        GemmArgs g;
        AttnArgs a;
        NormOpArgs n;
        ConvposOpArgs c;
        CfgOpArgs f;
    };
};
struct EvalEntry {               // solver behaviour after evaluation e (torchdiffeq Euler / Midpoint step)
    float coef;                  // x_new = x_state + coef * v
    int write_state;             // 1: x_state = x_new (end of a step); 0: x_new only feeds the next evaluation (midpoint's first half)
};
struct MegaArgs {
    const MegaOp* ops;           // device array, one network evaluation
    int n_ops;
    const EvalEntry* evals;      // device array
    int n_evals;
    unsigned* barrier;           // zeroed before the launch
    int* abort_flag;             // set if a grid barrier times out (never in a healthy run)
    long long* trace;            // optional: [n_ops][2] clock64 of CTA 0 (work done, barrier passed) during evaluation 1; or null
};

constexpr int MEGA_THREADS = 384;
constexpr int MEGA_WORKER0 = 4;                  // first worker warp
constexpr int MEGA_WORKERS = 256;                // worker threads
constexpr int MEGA_SMEM_BYTES = GemmCfg<128>::SMEM_BYTES > ATT_SMEM_BYTES ? GemmCfg<128>::SMEM_BYTES : ATT_SMEM_BYTES;
static_assert(GemmCfg<256>::SMEM_BYTES <= MEGA_SMEM_BYTES && GemmCfg<64>::SMEM_BYTES <= MEGA_SMEM_BYTES, "GEMM stages must fit");

__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }

// AdaptiveRMSNorm / RMSNorm rows (kernels.cuh rmsnorm_kernel), one worker warp per row, L2 loads.
template <int V>
__device__ __forceinline__ void mega_norm(const NormOpArgs& a, int eval, int cta, int n_ctas) {
    constexpr int D = 128 * V;
    const int wwarp = (threadIdx.x >> 5) - MEGA_WORKER0;
    const int lane = threadIdx.x & 31;
    if (wwarp < 0) return;
    const float* gamma = a.gamma + static_cast<long long>(eval) * a.gb_stride;
    const float* beta = a.beta ? a.beta + static_cast<long long>(eval) * a.gb_stride : nullptr;
    for (int row = cta * 8 + wwarp; row < a.M; row += n_ctas * 8) {
        const float4* xr = reinterpret_cast<const float4*>(a.x + static_cast<size_t>(row) * D);
        float4 v[V];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i] = __ldcg(xr + i * 32 + lane);
            ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float scale = sqrtf(static_cast<float>(D)) / fmaxf(sqrtf(ss), 1e-12f);
        uint2* orow = reinterpret_cast<uint2*>(a.out + static_cast<size_t>(row) * D);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
            float4 bt = make_float4(0.f, 0.f, 0.f, 0.f);
            if (beta != nullptr) bt = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i].x * scale * g.x + bt.x, v[i].y * scale * g.y + bt.y);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i].z * scale * g.z + bt.z, v[i].w * scale * g.w + bt.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            orow[i * 32 + lane] = pk;
        }
    }
}

// ConvPositionEmbed + residual (kernels.cuh convpos_kernel<31, TB>): virtual blocks of 256 channels x TB positions.
template <int KS, int TB>
__device__ __forceinline__ void mega_convpos(const ConvposOpArgs& a, int cta, int n_ctas) {
    const int tid = static_cast<int>(threadIdx.x) - MEGA_WORKER0 * 32;
    if (tid < 0) return;
    const int gx = (a.D + 255) / 256, gy = (a.N + TB - 1) / TB;
    const int total = gx * gy * a.Bt;
    for (int vb = cta; vb < total; vb += n_ctas) {
        const int c = (vb % gx) * 256 + tid;
        const int n0 = ((vb / gx) % gy) * TB;
        const int b = vb / (gx * gy);
        if (c >= a.D) continue;
        const float* hb = a.h + static_cast<size_t>(b) * a.N * a.D + c;
        float w[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) w[k] = __ldg(a.wT + k * a.D + c);
        float win[KS + TB - 1];
#pragma unroll
        for (int i = 0; i < KS + TB - 1; ++i) {
            const int n = n0 + i - KS / 2;
            win[i] = (n >= 0 && n < a.N) ? __ldcg(hb + static_cast<size_t>(n) * a.D) : 0.f;
        }
        const float bs = __ldg(a.bias + c);
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const int n = n0 + t;
            if (n >= a.N) break;
            float acc = bs;
#pragma unroll
            for (int k = 0; k < KS; ++k) acc = fmaf(w[k], win[t + k], acc);
            const float v = gelu_erf(acc) + win[t + KS / 2];
            const size_t off = (static_cast<size_t>(b) * a.N + n) * a.D + c;
            a.x[off] = v;
            a.x_h[off] = __float2bfloat16(v);
        }
    }
}

// CFG combine + solver update + re-quantised network input (kernels.cuh cfg_update_kernel).
__device__ __forceinline__ void mega_cfg(const CfgOpArgs& a, const EvalEntry ev, int cta, int n_ctas) {
    const int tid = static_cast<int>(threadIdx.x) - MEGA_WORKER0 * 32;
    if (tid < 0) return;
    const int total = a.BN * a.ldx;
    for (int gidx = cta * MEGA_WORKERS + tid; gidx < total; gidx += n_ctas * MEGA_WORKERS) {
        const int r = gidx / a.ldx, c = gidx % a.ldx;
        if (c >= a.dx) {
            a.xin[static_cast<size_t>(r) * a.ldx + c] = __float2bfloat16(0.f);
            if (a.two_branch) a.xin[static_cast<size_t>(a.BN + r) * a.ldx + c] = __float2bfloat16(0.f);
            continue;
        }
        const int idx = r * a.dx + c;
        float v = __ldcg(a.vpred + idx);
        if (a.two_branch) v = (1.0f + a.s) * v - a.s * __ldcg(a.vpred + static_cast<size_t>(a.BN) * a.dx + idx);
        const float xn = __ldcg(a.x_state + idx) + ev.coef * v;
        if (ev.write_state) a.x_state[idx] = xn;
        const __nv_bfloat16 hb = __float2bfloat16(xn);
        a.xin[static_cast<size_t>(r) * a.ldx + c] = hb;
        if (a.two_branch) a.xin[static_cast<size_t>(a.BN + r) * a.ldx + c] = hb;
    }
}

// Grid barrier (all CTAs co-resident: cooperative launch).  Returns true if it timed out (abort).
__device__ __forceinline__ bool mega_grid_barrier(const MegaArgs& m, unsigned& epoch, int* s_abort) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(m.barrier), "r"(1u) : "memory");
        const long long t0 = clock64();
        unsigned v;
        int ab = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(m.barrier) : "memory");
            if (v < epoch && clock64() - t0 > 2000000000ll) {          // ~1 s: a peer is gone
                *m.abort_flag = 1;
                ab = 1;
                break;
            }
        } while (v < epoch);
        if (!ab && *reinterpret_cast<volatile int*>(m.abort_flag)) ab = 1;
        *s_abort = ab;
    }
    __syncthreads();
    // data written by other CTAs through the generic proxy is read by TMA (async proxy) next
    asm volatile("fence.proxy.async;" ::: "memory");
    return *s_abort != 0;
}

// The op loop of one role: ROLE 1 = control warps 0-3 (TMA producer, MMA issuer), ROLE 2 = worker warps 4-11 (epilogue,
// softmax, element-wise).  Both loops execute the same sequence of CTA-wide synchronisations.
template <int POLY_MASK, int V, int ROLE>
__device__ __forceinline__ void mega_loop(const MegaArgs& margs, MegaOp& s_op, int* s_abort, uint8_t* smem, uint32_t tmem_base) {
    const int cta = blockIdx.x, n_ctas = gridDim.x;
    unsigned epoch = 0;
    for (int e = 0; e < margs.n_evals; ++e) {
        for (int i = 0; i < margs.n_ops; ++i) {
            if (ROLE == 2) {   // the op's scalar fields into shared memory (its TMA descriptors are used in place, from global memory)
                const uint4* src = reinterpret_cast<const uint4*>(margs.ops + i);
                uint4* dst = reinterpret_cast<uint4*>(&s_op);
                for (int k = threadIdx.x - MEGA_WORKER0 * 32; k < static_cast<int>(sizeof(MegaOp) / 16); k += MEGA_WORKERS) dst[k] = __ldg(src + k);
            }
            __syncthreads();
            const MegaOp* gop = margs.ops + i;
            switch (s_op.type) {
                case MOP_GEMM:
                    if (s_op.bn == 256) gemm_run<256, 1, ROLE>(&gop->g, s_op.g, smem, tmem_base, MEGA_WORKER0, cta, n_ctas);
                    else if (s_op.bn == 128) gemm_run<128, 1, ROLE>(&gop->g, s_op.g, smem, tmem_base, MEGA_WORKER0, cta, n_ctas);
                    else gemm_run<64, 1, ROLE>(&gop->g, s_op.g, smem, tmem_base, MEGA_WORKER0, cta, n_ctas);
                    tc_fence_before();          // the next op reuses the TMEM columns
                    break;
                case MOP_ATTN:
                    attention_run<POLY_MASK, 0, ROLE, false>(&gop->a, s_op.a, smem, tmem_base, cta, n_ctas);
                    tc_fence_before();
                    break;
                case MOP_NORM:
                    if (ROLE == 2) mega_norm<V>(s_op.n, e, cta, n_ctas);
                    break;
                case MOP_CONVPOS:
                    if (ROLE == 2) mega_convpos<31, 32>(s_op.c, cta, n_ctas);
                    break;
                default:
                    if (ROLE == 2) mega_cfg(s_op.f, margs.evals[e], cta, n_ctas);
                    break;
            }
            const bool tr = margs.trace != nullptr && e == 1 && cta == 0 && threadIdx.x == (ROLE == 2 ? MEGA_WORKER0 * 32 : 0);
            if (tr && ROLE == 2) margs.trace[2 * i] = clock64();
            if (mega_grid_barrier(margs, epoch, s_abort)) return;
            if (tr && ROLE == 1) margs.trace[2 * i + 1] = clock64();
            tc_fence_after();
        }
    }
}

template <int POLY_MASK, int V>
__global__ void __launch_bounds__(MEGA_THREADS, 1) flow_persistent_kernel(const __grid_constant__ MegaArgs margs) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ MegaOp s_op;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_abort;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) {
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // one split of the warps for the whole kernel, so that each role's code is compiled against its own register budget
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        mega_loop<POLY_MASK, V, 1>(margs, s_op, &s_abort, smem, tmem_base);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        mega_loop<POLY_MASK, V, 2>(margs, s_op, &s_abort, smem, tmem_base);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace covo
