// Flow-matching acoustic decoder: host-side plan (buffers, GEMM ops, CUDA graph) for
// ConditionalFlowMatcherWrapper.sample / CoVoMix.forward_with_cond_scale (covomix/covomix_model/acoustic.py).
#pragma once
#include "common.cuh"
#include "flow_persistent.cuh"

namespace covo {

constexpr int FLOW_MAX_TIMES = 128;
struct TimesArg {
    float t[FLOW_MAX_TIMES];
};
__global__ void set_times_kernel(float* out, TimesArg ta, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ta.t[i];
}

struct FlowLayerW {
    Tensor skip_w, skip_b;          // layers >= depth/2
    Tensor qkv_w, out_w, ff1_w, ff1_b, ff2_w, ff2_b;
};

struct FlowPlan {
    // key
    int B = 0, N = 0, method = 0, n_steps = 0, single_eval = 0;
    float cond_scale = 0.f;
    float step_size = 0.f;         // 0: uniform grid k / n_steps
    void* ws = nullptr;
    // derived
    int two_branch = 1, M = 0, BN = 0, n_t = 0;
    float times[FLOW_MAX_TIMES];
    float dts[FLOW_MAX_TIMES];
    // workspace buffers
    long long* ids = nullptr;
    float *cond = nullptr, *x_state = nullptr, *x_in = nullptr, *v_out = nullptr;
    float *d_times = nullptr, *tfeat = nullptr, *temb = nullptr, *gb = nullptr;
    float2* rope = nullptr;
    __nv_bfloat16 *a_pc = nullptr, *xin = nullptr, *slots = nullptr, *a_norm = nullptr, *qkv = nullptr, *attn_o = nullptr,
                  *ffh = nullptr;
    float *e_const = nullptr, *h0 = nullptr, *x = nullptr, *vpred = nullptr;
    // Tables that depend on the evaluation times and the weights only (time MLP, AdaLN gamma/beta, RoPE): plan-owned device memory, filled once when the plan is built (per call for covo_flow_velocity,
    // whose t is an argument) -- never in the caller's workspace, which may be recycled between calls.
    void* tables = nullptr;
    // ops
    GemmOp op_const, op_embed, op_pred;
    std::vector<GemmOp> op_skip, op_qkv, op_out, op_ff1, op_ff2;
    AttnArgs attn;
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
    // persistent flow-step kernel (flow_persistent.cuh): op list of one evaluation + per-evaluation solver table, owned by the plan
    bool persistent = false;
    MegaOp* d_ops = nullptr;
    EvalEntry* d_evals = nullptr;
    unsigned* d_sync = nullptr;      // [0] grid barrier counter, [1] abort flag
    int n_ops = 0;
    double flops_per_eval = 0.0;
};

}  // namespace covo

struct covo_flow {
    covo_flow_cfg cfg;
    covo::DeviceInfo di;
    covo::Weights w;
    // weight views
    covo::Tensor null_cond, time_w, time_lin_w, time_lin_b, emb_table, embed_wx, embed_wpc, embed_b, conv_wT, conv_b, inv_freq,
        gb_w, gb_b, final_gamma, pred_w;
    std::vector<covo::FlowLayerW> layers;
    int kpc = 0;      // padded width of [emb | cond]
    int ldx = 0;      // padded width of the x operand (dim_x -> multiple of 64)
    int npred = 0;    // padded rows of to_pred
    std::vector<covo::FlowPlan*> plans;
    cudaStream_t capture_stream = nullptr;
    bool use_graph = true;
    bool naive_attn = false;
    int last_launches = 0;         // kernels launched by the most recent covo_flow_sample call
    int persistent_mode = 0;       // COVO_FLOW_PERSISTENT: 0 (default) never, 1 whenever supported, -1 for M <= persistent_max_rows
    int persistent_max_rows = 4096;
    int persistent_bn = 0;         // COVO_FLOW_PERSISTENT_BN: force the GEMM tile width of the persistent path (0: heuristic)
    float step_size = 0.f;         // covo_flow_set_step_size: torchdiffeq grid k*h with the last point snapped to 1
    int serpentine = 1;            // COVO_FLOW_SERPENTINE: alternate the row order from kernel to kernel (flow_enqueue_network)
};

namespace covo {

inline int flow_bind_weights(covo_flow* h) {
    const covo_flow_cfg& c = h->cfg;
    const Weights& w = h->w;
    COVO_TRY(w.get("null_cond", DT_F32, &h->null_cond));
    COVO_TRY(w.get("time.w", DT_F32, &h->time_w));
    COVO_TRY(w.get("time.lin.w", DT_F32, &h->time_lin_w));
    COVO_TRY(w.get("time.lin.b", DT_F32, &h->time_lin_b));
    COVO_TRY(w.get("emb.table", DT_F32, &h->emb_table));
    COVO_TRY(w.get("embed.wx", DT_BF16, &h->embed_wx));
    COVO_TRY(w.get("embed.wpc", DT_BF16, &h->embed_wpc));
    COVO_TRY(w.get("embed.b", DT_F32, &h->embed_b));
    COVO_TRY(w.get("convpos.wT", DT_F32, &h->conv_wT));
    COVO_TRY(w.get("convpos.b", DT_F32, &h->conv_b));
    COVO_TRY(w.get("rope.inv_freq", DT_F32, &h->inv_freq));
    COVO_TRY(w.get("adaln.w", DT_F32, &h->gb_w));
    COVO_TRY(w.get("adaln.b", DT_F32, &h->gb_b));
    COVO_TRY(w.get("final.gamma", DT_F32, &h->final_gamma));
    COVO_TRY(w.get("pred.w", DT_BF16, &h->pred_w));
    h->ldx = static_cast<int>(h->embed_wx.shape[1]);
    h->kpc = static_cast<int>(h->embed_wpc.shape[1]);
    h->npred = static_cast<int>(h->pred_w.shape[0]);
    if (h->ldx % 64 || h->kpc % 64 || h->npred % 64 || h->ldx < c.dim_x ||
        h->kpc < c.n_streams * c.dim_phoneme_emb + c.dim_in)
        return fail(COVO_ERR_WEIGHTS, "packed embed/pred weights have unexpected padding (%d, %d, %d)", h->ldx, h->kpc, h->npred);
    h->layers.resize(c.depth);
    for (int L = 0; L < c.depth; ++L) {
        FlowLayerW& lw = h->layers[L];
        const std::string p = "L" + std::to_string(L) + ".";
        if (L >= c.depth / 2) {
            COVO_TRY(w.get(p + "skip.w", DT_BF16, &lw.skip_w));
            COVO_TRY(w.get(p + "skip.b", DT_F32, &lw.skip_b));
        }
        COVO_TRY(w.get(p + "qkv.w", DT_BF16, &lw.qkv_w));
        COVO_TRY(w.get(p + "out.w", DT_BF16, &lw.out_w));
        COVO_TRY(w.get(p + "ff1.w", DT_BF16, &lw.ff1_w));
        COVO_TRY(w.get(p + "ff1.b", DT_F32, &lw.ff1_b));
        COVO_TRY(w.get(p + "ff2.w", DT_BF16, &lw.ff2_w));
        COVO_TRY(w.get(p + "ff2.b", DT_F32, &lw.ff2_b));
    }
    return COVO_OK;
}

inline int flow_num_times(int method, int n_steps) { return method == COVO_ODE_MIDPOINT ? 2 * n_steps : n_steps; }

// Lays the plan's buffers out in the workspace (p.ws may be null: size query only).  Returns bytes used.
inline size_t flow_layout(const covo_flow* h, FlowPlan& p) {
    const covo_flow_cfg& c = h->cfg;
    const int D = c.dim, inner = c.heads * c.dim_head, M = p.M, BN = p.BN;
    Arena a(p.ws, static_cast<size_t>(-1));
    p.ids = a.take<long long>(static_cast<size_t>(BN) * c.n_streams);
    p.cond = a.take<float>(static_cast<size_t>(BN) * c.dim_in);
    p.x_state = a.take<float>(static_cast<size_t>(BN) * c.dim_x);
    p.x_in = a.take<float>(static_cast<size_t>(BN) * c.dim_x);
    p.v_out = a.take<float>(static_cast<size_t>(BN) * c.dim_x);
    p.a_pc = a.take<__nv_bfloat16>(static_cast<size_t>(M) * h->kpc);
    p.xin = a.take<__nv_bfloat16>(static_cast<size_t>(M) * h->ldx);
    p.slots = a.take<__nv_bfloat16>(static_cast<size_t>(c.depth / 2 + 1) * M * D);
    p.a_norm = a.take<__nv_bfloat16>(static_cast<size_t>(M) * D);
    p.qkv = a.take<__nv_bfloat16>(static_cast<size_t>(M) * 3 * inner);
    p.attn_o = a.take<__nv_bfloat16>(static_cast<size_t>(M) * inner);
    p.ffh = a.take<__nv_bfloat16>(static_cast<size_t>(M) * D * c.ff_mult);
    p.e_const = a.take<float>(static_cast<size_t>(M) * D);
    p.h0 = a.take<float>(static_cast<size_t>(M) * D);
    p.x = a.take<float>(static_cast<size_t>(M) * D);
    p.vpred = a.take<float>(static_cast<size_t>(M) * c.dim_x);
    return align_up(a.off, 256);
}

// Plan-owned tables (see FlowPlan::tables).  base == null: size query.
inline size_t flow_layout_tables(const covo_flow* h, FlowPlan& p, void* base) {
    const covo_flow_cfg& c = h->cfg;
    const int D = c.dim;
    Arena a(base, static_cast<size_t>(-1));
    p.d_times = a.take<float>(FLOW_MAX_TIMES);
    p.tfeat = a.take<float>(static_cast<size_t>(p.n_t) * D);
    p.temb = a.take<float>(static_cast<size_t>(p.n_t) * D * 4);
    p.gb = a.take<float>(static_cast<size_t>(p.n_t) * c.depth * 4 * D);
    p.rope = a.take<float2>(static_cast<size_t>(p.N) * 32);
    return align_up(a.off, 256);
}

inline void flow_times(FlowPlan& p) {
    // torchdiffeq fixed grid: grid[k] = k*h (+t0 = 0), last point snapped to 1; dt = grid[k+1]-grid[k];
    // midpoint evaluates at grid[k] and grid[k] + dt/2.  All in fp32 like the reference.
    // (step sizes that do not divide 1, e.g. 0.3, give 0, .3, .6, .9, 1: the last step is shorter.)
    const float hstep = p.step_size > 0.f ? p.step_size : 1.0f / static_cast<float>(p.n_steps);
    int idx = 0;
    for (int k = 0; k < p.n_steps; ++k) {
        const float t0 = static_cast<float>(k) * hstep;
        const float t1 = (k + 1 == p.n_steps) ? 1.0f : static_cast<float>(k + 1) * hstep;
        const float dt = t1 - t0;
        p.times[idx] = t0;
        p.dts[idx] = dt;
        ++idx;
        if (p.method == COVO_ODE_MIDPOINT) {
            p.times[idx] = t0 + 0.5f * dt;
            p.dts[idx] = dt;
            ++idx;
        }
    }
}

inline ASource a2d(const void* ptr, int K, int rows) {
    ASource a;
    a.ptr = ptr;
    a.K = K;
    a.rows = rows;
    a.Z = 1;
    a.row_stride = K;
    a.z_stride = static_cast<long long>(K) * rows;
    return a;
}

inline int flow_build_ops(covo_flow* h, FlowPlan& p) {
    const covo_flow_cfg& c = h->cfg;
    const int D = c.dim, inner = c.heads * c.dim_head, M = p.M, half = c.depth / 2;
    const int fbn = p.persistent ? h->persistent_bn : 0;       // tile-width override of the persistent path (experiments)
    // per-call constant part of to_embed:  e_const = [emb | cond] W_pc^T + b
    gemm_defaults(p.op_const.args);
    COVO_TRY(build_gemm(p.op_const, h->di, a2d(p.a_pc, h->kpc, M), M, 1, h->embed_wpc.ptr, D, 1, 0));
    COVO_TRY(gemm_set_outputs(p.op_const, p.e_const, nullptr, nullptr, D, M, 1, D, 0, 0));
    p.op_const.args.bias = h->embed_b.as<float>();
    p.op_const.flops = 2.0 * M * D * (c.n_streams * c.dim_phoneme_emb + c.dim_in);
    // per-evaluation part: h0 = x W_x^T + e_const
    gemm_defaults(p.op_embed.args);
    COVO_TRY(build_gemm(p.op_embed, h->di, a2d(p.xin, h->ldx, M), M, 1, h->embed_wx.ptr, D, 1, 0, fbn));
    COVO_TRY(gemm_set_outputs(p.op_embed, p.h0, p.e_const, nullptr, D, M, 1, D, 0, 0));
    p.op_embed.flops = 2.0 * M * D * c.dim_x;

    p.op_skip.assign(c.depth, GemmOp());
    p.op_qkv.assign(c.depth, GemmOp());
    p.op_out.assign(c.depth, GemmOp());
    p.op_ff1.assign(c.depth, GemmOp());
    p.op_ff2.assign(c.depth, GemmOp());
    for (int L = 0; L < c.depth; ++L) {
        const FlowLayerW& lw = h->layers[L];
        if (L >= half) {
            // x = Linear(cat(x, skip)): two taps over the bf16 slot tensor [slots][M][D]
            GemmOp& op = p.op_skip[L];
            gemm_defaults(op.args);
            ASource a;
            a.ptr = p.slots;
            a.K = D;
            a.rows = M;
            a.Z = half + 1;
            a.row_stride = D;
            a.z_stride = static_cast<long long>(M) * D;
            COVO_TRY(build_gemm(op, h->di, a, M, 1, lw.skip_w.ptr, D, 2, 0, fbn));
            op.args.tap_row[0] = op.args.tap_row[1] = 0;
            op.args.tap_z[0] = half;                    // current x
            op.args.tap_z[1] = c.depth - 1 - L;         // LIFO pop (acoustic.py:306-310)
            COVO_TRY(gemm_set_outputs(op, p.x, nullptr, nullptr, D, M, 1, D, 0, 0));
            op.args.bias = lw.skip_b.as<float>();
            op.flops = 2.0 * M * D * 2 * D;
        }
        {
            GemmOp& op = p.op_qkv[L];
            gemm_defaults(op.args);
            COVO_TRY(build_gemm(op, h->di, a2d(p.a_norm, D, M), M, 1, lw.qkv_w.ptr, 3 * inner, 1, 0, fbn));
            COVO_TRY(gemm_set_outputs(op, nullptr, nullptr, p.qkv, 3 * inner, M, 1, 3 * inner, 0, 0));
            op.args.rope = p.rope;
            op.args.rope_seq = p.N;
            op.args.rope_cols = 2 * inner;
            op.flops = 2.0 * M * 3 * inner * D;
        }
        {
            GemmOp& op = p.op_out[L];
            gemm_defaults(op.args);
            COVO_TRY(build_gemm(op, h->di, a2d(p.attn_o, inner, M), M, 1, lw.out_w.ptr, D, 1, 0, fbn));
            COVO_TRY(gemm_set_outputs(op, p.x, p.x, nullptr, D, M, 1, D, 0, 0));
            op.flops = 2.0 * M * D * inner;
        }
        {
            GemmOp& op = p.op_ff1[L];
            gemm_defaults(op.args);
            COVO_TRY(build_gemm(op, h->di, a2d(p.a_norm, D, M), M, 1, lw.ff1_w.ptr, D * c.ff_mult, 1, 0, fbn));
            COVO_TRY(gemm_set_outputs(op, nullptr, nullptr, p.ffh, D * c.ff_mult, M, 1, D * c.ff_mult, 0, 0));
            op.args.bias = lw.ff1_b.as<float>();
            op.args.act_h = ACT_GELU;
            op.flops = 2.0 * M * D * c.ff_mult * D;
        }
        {
            GemmOp& op = p.op_ff2[L];
            gemm_defaults(op.args);
            COVO_TRY(build_gemm(op, h->di, a2d(p.ffh, D * c.ff_mult, M), M, 1, lw.ff2_w.ptr, D, 1, 0, fbn));
            // bf16 copy of the next layer's input: a skip slot (first half) or the "current" slot (second half)
            __nv_bfloat16* next_h = nullptr;
            if (L + 1 < c.depth) {
                const int slot = (L + 1 < half) ? (L + 1) : half;
                next_h = p.slots + static_cast<size_t>(slot) * M * D;
            }
            COVO_TRY(gemm_set_outputs(op, p.x, p.x, next_h, D, M, 1, D, 0, 0));
            op.args.bias = lw.ff2_b.as<float>();

            op.flops = 2.0 * M * D * c.ff_mult * D;
        }
    }
    gemm_defaults(p.op_pred.args);
    COVO_TRY(build_gemm(p.op_pred, h->di, a2d(p.a_norm, D, M), M, 1, h->pred_w.ptr, h->npred, 1, 0, fbn));
    COVO_TRY(gemm_set_outputs(p.op_pred, p.vpred, nullptr, nullptr, c.dim_x, M, 1, c.dim_x, 0, 0));
    p.op_pred.flops = 2.0 * M * D * c.dim_x;

    // attention
    if (c.dim_head != ATT_D) return fail(COVO_ERR_INVALID, "dim_head=%d unsupported (64 only)", c.dim_head);
    COVO_TRY(attn_build_args(p.attn, p.qkv, p.attn_o, M / p.N, p.N, c.heads));
    return COVO_OK;
}

inline int launch_attention(const covo_flow* h, const FlowPlan& p, cudaStream_t st, int rev = 0) {
    const int Bt = p.M / p.N;
    ProfScope ps(PC_ATTN, 4.0 * p.N * static_cast<double>(p.N) * h->cfg.dim_head * h->cfg.heads * Bt, st);
    if (h->naive_attn) {
        dim3 grid(ceil_div(p.N, 8), h->cfg.heads, Bt);
        naive_attention_kernel<<<grid, 256, 0, st>>>(p.qkv, p.attn_o, p.N, h->cfg.heads, p.attn.inner,
                                                     1.0f / sqrtf(static_cast<float>(h->cfg.dim_head)));
    } else {
        AttnArgs a = p.attn;
        a.reverse = rev;
        COVO_TRY(launch_attention_kernel(a, h->di.num_sms, st));
    }
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

template <int V>
inline void launch_rmsnorm_v(const float* x, const float* g, const float* b, __nv_bfloat16* out, int M, int rev, cudaStream_t st) {
    rmsnorm_kernel<V><<<ceil_div(M, 8), 256, 0, st>>>(x, g, b, out, M, rev);
}
inline int launch_rmsnorm(const float* x, const float* g, const float* b, __nv_bfloat16* out, int M, int D, int rev, cudaStream_t st) {
    ProfScope ps(PC_NORM, 0.0, st);
    switch (D / 128) {
        case 8: launch_rmsnorm_v<8>(x, g, b, out, M, rev, st); break;
        case 4: launch_rmsnorm_v<4>(x, g, b, out, M, rev, st); break;
        case 2: launch_rmsnorm_v<2>(x, g, b, out, M, rev, st); break;
        case 1: launch_rmsnorm_v<1>(x, g, b, out, M, rev, st); break;
        default: return fail(COVO_ERR_INVALID, "dim %d unsupported by rmsnorm (need 128/256/512/1024)", D);
    }
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

// Tables that depend on the evaluation times only: time MLP (acoustic.py:361-365), the 4 x depth AdaLN projections of every
// time (:198-204), RoPE factors (:126-130).  Once per plan (per call for single_eval).
inline int flow_enqueue_tables(covo_flow* h, FlowPlan& p, cudaStream_t st, int* launches) {
    const covo_flow_cfg& c = h->cfg;
    const int D = c.dim;
    TimesArg ta;
    ProfScope ps(PC_PROLOGUE, 0.0, st);
    memcpy(ta.t, p.times, sizeof(float) * p.n_t);
    set_times_kernel<<<1, FLOW_MAX_TIMES, 0, st>>>(p.d_times, ta, p.n_t);
    time_features_kernel<<<p.n_t, 256, 0, st>>>(p.d_times, h->time_w.as<float>(), p.tfeat, p.n_t, D / 2);
    {
        dim3 g(ceil_div(4 * D, 64), ceil_div(p.n_t, 64));
        sgemm_nt_kernel<float><<<g, 256, 0, st>>>(p.tfeat, D, h->time_lin_w.as<float>(), h->time_lin_b.as<float>(), p.temb, 4 * D,
                                                  p.n_t, 4 * D, D, SG_SILU);
    }
    const int NO = c.depth * 4 * D;           // per layer: gamma1 | beta1 | gamma2 | beta2
    {
        dim3 g(ceil_div(NO, 64), ceil_div(p.n_t, 64));
        sgemm_nt_kernel<float><<<g, 256, 0, st>>>(p.temb, 4 * D, h->gb_w.as<float>(), h->gb_b.as<float>(), p.gb, NO, p.n_t, NO, 4 * D,
                                                  SG_NONE);
    }
    *launches += 4;
    rope_table_kernel<<<ceil_div(p.N * 32, 256), 256, 0, st>>>(h->inv_freq.as<float>(), p.rope, p.N, 32);
    ++*launches;
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

// Work that depends on the call's inputs but not on the ODE state: e_const = [emb | cond] W_pc^T + b.
inline int flow_enqueue_prologue(covo_flow* h, FlowPlan& p, cudaStream_t st, int* launches) {
    const covo_flow_cfg& c = h->cfg;
    if (p.single_eval) COVO_TRY(flow_enqueue_tables(h, p, st, launches));
    {
        ProfScope ps(PC_PROLOGUE, 0.0, st);
        embed_input_kernel<<<p.M, 256, 0, st>>>(p.ids, p.cond, h->emb_table.as<float>(), h->null_cond.as<float>(), p.a_pc, p.BN,
                                                c.n_streams, c.dim_phoneme_emb, c.dim_in, c.num_phoneme_tokens, h->kpc);
    }
    COVO_CK(cudaGetLastError());
    COVO_TRY(launch_gemm(p.op_const, st));
    *launches += 2;
    return COVO_OK;
}

// One network pass over both CFG branches: xin (bf16 state) -> vpred [M, dim_x].
inline int flow_enqueue_network(covo_flow* h, FlowPlan& p, int t_idx, cudaStream_t st, int* launches) {
    const covo_flow_cfg& c = h->cfg;
    const int D = c.dim, M = p.M, half = c.depth / 2;
    const int Bt = M / p.N;
    COVO_TRY(launch_gemm(p.op_embed, st));
    {
        ProfScope ps(PC_CONVPOS, 0.0, st);
        // 32 positions per thread: 62 window loads per 32 outputs (1.9x read amplification; 8 per thread was 4.75x)
        constexpr int CONVPOS_TB = 32;
        dim3 g(ceil_div(D, 256), ceil_div(p.N, CONVPOS_TB), Bt);
        convpos_kernel<31, CONVPOS_TB><<<g, 256, 0, st>>>(p.h0, h->conv_wT.as<float>(), h->conv_b.as<float>(), p.x, p.slots, p.N, D,
                                                          h->serpentine);
        COVO_CK(cudaGetLastError());
    }
    *launches += 2;
    const float* gb_t = p.gb + static_cast<size_t>(t_idx) * c.depth * 4 * D;
    // Serpentine row order along the chain (COVO_FLOW_SERPENTINE, default on): every kernel walks the token rows in the
    // opposite direction of its producer, so it starts on the rows that were written LAST and are still in L2.  The
    // activations of a C3 pass (x 108 MB fp32, a_norm 54, qkv 162, ffh 216 MB) are each about the size of the 126 MB L2: in
    // producer order a consumer finds its first rows evicted already, LRU-thrashing through the whole tensor.
    int dir = h->serpentine;                          // to_embed ran ascending, the conv-pos kernel descending
    auto turn = [&]() { dir = h->serpentine ? dir ^ 1 : 0; return dir; };
    auto gemm_dir = [&](const GemmOp& op) -> int {
        GemmOp o = op;
        o.args.reverse = turn();
        return launch_gemm(o, st);
    };
    for (int L = 0; L < c.depth; ++L) {
        const float* gbl = gb_t + static_cast<size_t>(L) * 4 * D;
        if (L >= half) {
            COVO_TRY(gemm_dir(p.op_skip[L]));
            ++*launches;
        }
        COVO_TRY(launch_rmsnorm(p.x, gbl, gbl + D, p.a_norm, M, D, turn(), st));
        COVO_TRY(gemm_dir(p.op_qkv[L]));
        COVO_TRY(launch_attention(h, p, st, turn()));
        COVO_TRY(gemm_dir(p.op_out[L]));
        COVO_TRY(launch_rmsnorm(p.x, gbl + 2 * D, gbl + 3 * D, p.a_norm, M, D, turn(), st));
        COVO_TRY(gemm_dir(p.op_ff1[L]));
        COVO_TRY(gemm_dir(p.op_ff2[L]));
        *launches += 7;
    }
    COVO_TRY(launch_rmsnorm(p.x, h->final_gamma.as<float>(), nullptr, p.a_norm, M, D, turn(), st));
    COVO_TRY(gemm_dir(p.op_pred));
    *launches += 2;
    return COVO_OK;
}

inline int flow_launch_persistent(covo_flow* h, FlowPlan& p, cudaStream_t st);

inline int flow_enqueue_sample(covo_flow* h, FlowPlan& p, cudaStream_t st, int* launches) {
    const covo_flow_cfg& c = h->cfg;
    const int n_el = p.BN * h->ldx;                 // the element-wise kernels also rewrite the padding columns of xin
    const int eb = ceil_div(n_el, 256);
    COVO_TRY(flow_enqueue_prologue(h, p, st, launches));
    {
        ProfScope ps(PC_ELEMWISE, 0.0, st);
        state_to_input_kernel<<<eb, 256, 0, st>>>(p.x_state, p.xin, p.BN, c.dim_x, h->ldx, p.two_branch);
    }
    ++*launches;
    if (p.persistent) {
        COVO_TRY(flow_launch_persistent(h, p, st));      // every evaluation + solver update of the call: one launch
        ++*launches;
        return COVO_OK;
    }
    int ti = 0;
    for (int k = 0; k < p.n_steps; ++k) {
        if (p.method == COVO_ODE_MIDPOINT) {
            const float dt = p.dts[ti];
            COVO_TRY(flow_enqueue_network(h, p, ti, st, launches));
            // y_mid = y0 + f0 * dt/2 -> network input only
            {
                ProfScope ps(PC_ELEMWISE, 0.0, st);
                cfg_update_kernel<<<eb, 256, 0, st>>>(p.vpred, p.x_state, nullptr, nullptr, p.xin, p.BN, c.dim_x, h->ldx,
                                                  p.cond_scale, 0.5f * dt, p.two_branch);
            }
            ++ti;
            COVO_TRY(flow_enqueue_network(h, p, ti, st, launches));
            // y1 = y0 + dt * f(t0 + dt/2, y_mid)
            {
                ProfScope ps(PC_ELEMWISE, 0.0, st);
                cfg_update_kernel<<<eb, 256, 0, st>>>(p.vpred, p.x_state, p.x_state, nullptr, p.xin, p.BN, c.dim_x, h->ldx,
                                                  p.cond_scale, dt, p.two_branch);
            }
            ++ti;
            *launches += 2;
        } else {
            const float dt = p.dts[ti];
            COVO_TRY(flow_enqueue_network(h, p, ti, st, launches));
            {
                ProfScope ps(PC_ELEMWISE, 0.0, st);
                cfg_update_kernel<<<eb, 256, 0, st>>>(p.vpred, p.x_state, p.x_state, nullptr, p.xin, p.BN, c.dim_x, h->ldx,
                                                  p.cond_scale, dt, p.two_branch);
            }
            ++ti;
            ++*launches;
        }
    }
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

inline int flow_enqueue_velocity(covo_flow* h, FlowPlan& p, cudaStream_t st, int* launches) {
    const covo_flow_cfg& c = h->cfg;
    const int n_el = p.BN * h->ldx;                 // the element-wise kernels also rewrite the padding columns of xin
    const int eb = ceil_div(n_el, 256);
    COVO_TRY(flow_enqueue_prologue(h, p, st, launches));
    {
        ProfScope ps(PC_ELEMWISE, 0.0, st);
        state_to_input_kernel<<<eb, 256, 0, st>>>(p.x_in, p.xin, p.BN, c.dim_x, h->ldx, p.two_branch);
    }
    COVO_TRY(flow_enqueue_network(h, p, 0, st, launches));
    {
        ProfScope ps(PC_ELEMWISE, 0.0, st);
        cfg_update_kernel<<<eb, 256, 0, st>>>(p.vpred, nullptr, nullptr, p.v_out, nullptr, p.BN, c.dim_x, h->ldx, p.cond_scale,
                                          0.f, p.two_branch);
    }
    COVO_CK(cudaGetLastError());
    *launches += 2;
    return COVO_OK;
}

inline void flow_free_plan(FlowPlan* p) {
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->d_ops) cudaFree(p->d_ops);
    if (p->d_evals) cudaFree(p->d_evals);
    if (p->d_sync) cudaFree(p->d_sync);
    if (p->tables) cudaFree(p->tables);
    delete p;
}

inline bool flow_persistent_eligible(const covo_flow* h, const FlowPlan& p) {
    if (h->naive_attn || p.single_eval || h->cfg.dim != 1024 || h->cfg.conv_pos_kernel != 31) return false;
    if (h->persistent_mode == 0) return false;
    if (h->persistent_mode == 1) return true;
    return p.M <= h->persistent_max_rows;
}

// Op list of one network evaluation (the launch sequence of flow_enqueue_network + the CFG / solver update) and the
// solver table, copied to plan-owned device memory.
inline int flow_build_persistent(covo_flow* h, FlowPlan& p) {
    const covo_flow_cfg& c = h->cfg;
    const int D = c.dim, M = p.M, half = c.depth / 2;
    std::vector<MegaOp> ops;
    double flops = 0.0;
    auto gemm = [&](const GemmOp& g) {
        MegaOp m;
        memset(&m, 0, sizeof(m));
        m.type = MOP_GEMM;
        m.bn = g.bn;
        m.g = g.args;
        ops.push_back(m);
        flops += g.flops;
    };
    auto norm = [&](const float* gamma, const float* beta, long long stride) {
        MegaOp m;
        memset(&m, 0, sizeof(m));
        m.type = MOP_NORM;
        m.n.x = p.x;
        m.n.gamma = gamma;
        m.n.beta = beta;
        m.n.out = p.a_norm;
        m.n.M = M;
        m.n.gb_stride = stride;
        ops.push_back(m);
    };
    gemm(p.op_embed);
    {
        MegaOp m;
        memset(&m, 0, sizeof(m));
        m.type = MOP_CONVPOS;
        m.c.h = p.h0;
        m.c.wT = h->conv_wT.as<float>();
        m.c.bias = h->conv_b.as<float>();
        m.c.x = p.x;
        m.c.x_h = p.slots;
        m.c.N = p.N;
        m.c.D = D;
        m.c.Bt = M / p.N;
        ops.push_back(m);
    }
    const long long gb_stride = static_cast<long long>(c.depth) * 4 * D;
    for (int L = 0; L < c.depth; ++L) {
        const float* gbl = p.gb + static_cast<size_t>(L) * 4 * D;
        if (L >= half) gemm(p.op_skip[L]);
        norm(gbl, gbl + D, gb_stride);
        gemm(p.op_qkv[L]);
        {
            MegaOp m;
            memset(&m, 0, sizeof(m));
            m.type = MOP_ATTN;
            m.a = p.attn;
            ops.push_back(m);
            flops += 4.0 * p.N * static_cast<double>(p.N) * c.dim_head * c.heads * (M / p.N);
        }
        gemm(p.op_out[L]);
        norm(gbl + 2 * D, gbl + 3 * D, gb_stride);
        gemm(p.op_ff1[L]);
        gemm(p.op_ff2[L]);
    }
    norm(h->final_gamma.as<float>(), nullptr, 0);
    gemm(p.op_pred);
    {
        MegaOp m;
        memset(&m, 0, sizeof(m));
        m.type = MOP_CFG;
        m.f.vpred = p.vpred;
        m.f.x_state = p.x_state;
        m.f.xin = p.xin;
        m.f.BN = p.BN;
        m.f.dx = c.dim_x;
        m.f.ldx = h->ldx;
        m.f.s = p.cond_scale;
        m.f.two_branch = p.two_branch;
        ops.push_back(m);
    }
    std::vector<EvalEntry> evals(p.n_t);
    for (int e = 0; e < p.n_t; ++e) {
        if (p.method == COVO_ODE_MIDPOINT) {
            evals[e].coef = (e & 1) ? p.dts[e] : 0.5f * p.dts[e];     // y_mid = y0 + f0 dt/2 ; y1 = y0 + dt f(t0 + dt/2, y_mid)
            evals[e].write_state = e & 1;
        } else {
            evals[e].coef = p.dts[e];
            evals[e].write_state = 1;
        }
    }
    p.n_ops = static_cast<int>(ops.size());
    p.flops_per_eval = flops;
    COVO_CK(cudaMalloc(&p.d_ops, ops.size() * sizeof(MegaOp)));
    COVO_CK(cudaMalloc(&p.d_evals, evals.size() * sizeof(EvalEntry)));
    COVO_CK(cudaMalloc(&p.d_sync, 32768));
    COVO_CK(cudaMemcpy(p.d_ops, ops.data(), ops.size() * sizeof(MegaOp), cudaMemcpyHostToDevice));
    COVO_CK(cudaMemcpy(p.d_evals, evals.data(), evals.size() * sizeof(EvalEntry), cudaMemcpyHostToDevice));
    return COVO_OK;
}

inline int flow_launch_persistent(covo_flow* h, FlowPlan& p, cudaStream_t st) {
    auto kern = flow_persistent_kernel<0x88, 8>;
    static bool attr = false;
    if (!attr) {
        COVO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_SMEM_BYTES));
        int per_sm = 0;
        COVO_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MEGA_THREADS, MEGA_SMEM_BYTES));
        if (per_sm < 1) return fail(COVO_ERR_INVALID, "persistent flow kernel does not fit on an SM");
        attr = true;
    }
    COVO_CK(cudaMemsetAsync(p.d_sync, 0, 32768, st));
    const bool trace = getenv("COVO_FLOW_TRACE") != nullptr && p.n_ops <= 480;
    MegaArgs a;
    a.trace = trace ? reinterpret_cast<long long*>(p.d_sync + 64) : nullptr;
    a.ops = p.d_ops;
    a.n_ops = p.n_ops;
    a.evals = p.d_evals;
    a.n_evals = p.n_t;
    a.barrier = p.d_sync;
    a.abort_flag = reinterpret_cast<int*>(p.d_sync + 1);
    void* params[1] = {&a};
    ProfScope ps(PC_FLOW_PERSISTENT, p.flops_per_eval * p.n_t, st);
    if (trace) {
        long long* gt = reinterpret_cast<long long*>(p.d_sync + 2048);      // internal timelines of the first 200 gemm ops
        const int zero = 0;
        COVO_CK(cudaMemcpyToSymbolAsync(g_gemm_trace, &gt, sizeof(gt), 0, cudaMemcpyHostToDevice, st));
        COVO_CK(cudaMemcpyToSymbolAsync(g_gemm_trace_n, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, st));
    }
    COVO_CK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(h->di.num_sms), dim3(MEGA_THREADS), params,
                                        MEGA_SMEM_BYTES, st));
    if (trace) {                // debug only: synchronises and prints CTA 0's per-op timeline of evaluation 1
        std::vector<long long> t(2 * p.n_ops);
        std::vector<MegaOp> ops(p.n_ops);
        COVO_CK(cudaStreamSynchronize(st));
        COVO_CK(cudaMemcpy(t.data(), p.d_sync + 64, t.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        COVO_CK(cudaMemcpy(ops.data(), p.d_ops, ops.size() * sizeof(MegaOp), cudaMemcpyDeviceToHost));
        static const char* nm[5] = {"gemm", "attn", "norm", "convpos", "cfg"};
        for (int i = 1; i < p.n_ops; ++i)
            fprintf(stderr, "[flow trace] op %2d %-7s bn=%3d tiles=%4d: work %6lld cycles, barrier %6lld\n", i, nm[ops[i].type], ops[i].bn,
                    ops[i].type == MOP_GEMM ? ops[i].g.Z * ((ops[i].g.rows + 127) / 128) * ops[i].g.n_tiles : 0, t[2 * i] - t[2 * i - 1],
                    t[2 * i + 1] - t[2 * i]);
        fprintf(stderr, "[flow trace] evaluation 1: %lld cycles for ops 1..%d\n", t[2 * p.n_ops - 1] - t[1], p.n_ops - 1);
        std::vector<long long> g(8 * 200);
        COVO_CK(cudaMemcpy(g.data(), p.d_sync + 2048, g.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const int per_eval = 5 * h->cfg.depth + h->cfg.depth / 2 + 2;
        for (int k = per_eval; k < 2 * per_eval && k < 200; ++k) {        // the gemm ops of evaluation 1
            const long long* r = g.data() + 8 * k;
            fprintf(stderr, "[flow trace] gemm #%2d: init %5lld | first stage landed +%5lld | last MMA issued +%6lld | accumulator complete +%5lld | "
                    "stores issued +%5lld | stores complete +%5lld\n", k - per_eval, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5]);
        }
        long long* null_ptr = nullptr;
        COVO_CK(cudaMemcpyToSymbol(g_gemm_trace, &null_ptr, sizeof(null_ptr)));
    }
    return COVO_OK;
}

// Finds or builds the plan for this call shape.  single_eval: covo_flow_velocity (no graph; t varies per call).
inline int flow_get_plan(covo_flow* h, int B, int N, int method, int n_steps, float cond_scale, int single_eval, void* ws,
                         size_t ws_bytes, FlowPlan** out) {
    if (B < 1 || N < 1) return fail(COVO_ERR_INVALID, "B=%d N=%d must be positive", B, N);
    if (method != COVO_ODE_EULER && method != COVO_ODE_MIDPOINT) return fail(COVO_ERR_INVALID, "unknown ODE method %d", method);
    const int n_t = single_eval ? 1 : flow_num_times(method, n_steps);
    if (n_steps < 1 || n_t > FLOW_MAX_TIMES) return fail(COVO_ERR_INVALID, "n_steps=%d out of range", n_steps);
    const float step_size = single_eval ? 0.f : h->step_size;
    if (step_size > 0.f && n_steps != static_cast<int>(ceilf(1.0f / step_size)))
        return fail(COVO_ERR_INVALID, "n_steps=%d does not match ceil(1/step_size) = %d (covo_flow_set_step_size(%g))", n_steps,
                    static_cast<int>(ceilf(1.0f / step_size)), step_size);
    for (FlowPlan* q : h->plans) {
        if (q->B == B && q->N == N && q->method == method && q->n_steps == n_steps && q->cond_scale == cond_scale &&
            q->single_eval == single_eval && q->ws == ws && q->step_size == step_size) {
            *out = q;
            return COVO_OK;
        }
    }
    FlowPlan* p = new FlowPlan();
    p->B = B;
    p->N = N;
    p->method = method;
    p->n_steps = n_steps;
    p->cond_scale = cond_scale;
    p->single_eval = single_eval;
    p->step_size = step_size;
    p->ws = ws;
    p->two_branch = (cond_scale != 1.0f) ? 1 : 0;     // acoustic.py:423-424: cond_scale == 1 returns the cond branch only
    p->BN = B * N;
    p->M = p->BN * (p->two_branch ? 2 : 1);
    p->n_t = n_t;
    if (!single_eval) flow_times(*p);
    const size_t need = flow_layout(h, *p);
    if (need > ws_bytes || ws == nullptr) {
        delete p;
        return fail(COVO_ERR_INVALID, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    }
    p->persistent = flow_persistent_eligible(h, *p);
    {
        const size_t tb = flow_layout_tables(h, *p, nullptr);
        if (cudaMalloc(&p->tables, tb) != cudaSuccess) {
            delete p;
            return fail(COVO_ERR_CUDA, "cudaMalloc of %zu bytes (time tables) failed", tb);
        }
        flow_layout_tables(h, *p, p->tables);
    }
    const int saved_mc = h->di.gemm_mc, saved_cg = h->di.gemm_cg;
    if (p->persistent) h->di.gemm_mc = h->di.gemm_cg = 1;          // the cooperative kernel is not launched in clusters
    int rc = flow_build_ops(h, *p);
    h->di.gemm_mc = saved_mc;
    h->di.gemm_cg = saved_cg;
    if (rc == COVO_OK && !single_eval) {
        int tl = 0;
        rc = flow_enqueue_tables(h, *p, h->capture_stream, &tl);
        if (rc == COVO_OK && cudaStreamSynchronize(h->capture_stream) != cudaSuccess) rc = fail(COVO_ERR_CUDA, "time tables failed");
    }
    if (rc != COVO_OK) {
        flow_free_plan(p);
        return rc;
    }
    // padding columns of the bf16 operands (a_pc, xin) must be exact zeros (they meet zero weight columns, but
    // 0 * NaN != 0): embed_input_kernel / state_to_input_kernel / cfg_update_kernel rewrite them on EVERY call, so a
    // cached plan stays valid when the caller reuses or re-allocates the workspace between calls.
    if (p->persistent) {
        rc = flow_build_persistent(h, *p);
        if (rc != COVO_OK) {
            flow_free_plan(p);
            return rc;
        }
    }
    if (!single_eval && h->use_graph && !p->persistent) {
        cudaGraph_t graph = nullptr;
        int launches = 0;
        COVO_CK(cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeThreadLocal));
        rc = flow_enqueue_sample(h, *p, h->capture_stream, &launches);
        cudaError_t e = cudaStreamEndCapture(h->capture_stream, &graph);
        if (rc != COVO_OK || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            flow_free_plan(p);
            return rc != COVO_OK ? rc : fail(COVO_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(e));
        }
        e = cudaGraphInstantiate(&p->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            flow_free_plan(p);
            return fail(COVO_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        }
        p->launches = launches;
    }
    if (h->plans.size() >= 8) {
        flow_free_plan(h->plans.front());
        h->plans.erase(h->plans.begin());
    }
    h->plans.push_back(p);
    *out = p;
    return COVO_OK;
}

}  // namespace covo
