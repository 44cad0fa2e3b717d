// libcovomix_b200.so -- C ABI (include/covomix_b200.h) over the sm_100a kernels.
#include "flow.cuh"
#include "hifigan.cuh"
#include "t2s.cuh"
#include "frontend.cuh"

using namespace covo;

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int init_kernel_attrs() {
    COVO_TRY((set_gemm_attr<256, 1>()));
    COVO_TRY((set_gemm_attr<128, 1>()));
    COVO_TRY((set_gemm_attr<64, 1>()));
    COVO_TRY((set_gemm_attr<256, 0>()));
    COVO_TRY((set_gemm_attr<128, 0>()));
    COVO_TRY((set_gemm_attr<64, 0>()));
    COVO_TRY(attn_set_attrs());
    COVO_CK(cudaFuncSetAttribute(hifigan_fused_last_stage_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(HF_SMEM_BYTES)));
    COVO_CK(cudaFuncSetAttribute(hifigan_fused_last_stage_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(HF_SMEM_BYTES)));
    return COVO_OK;
}

bool env_flag(const char* name) {
    const char* v = getenv(name);
    return v && v[0] && v[0] != '0';
}

}  // namespace

extern "C" {

const char* covo_last_error(void) { return err_slot().c_str(); }
int covo_version(void) { return COVO_ABI_VERSION; }

// ====================================================================================== flow
int covo_flow_create(const covo_flow_cfg* cfg, const void* packed_weights, size_t bytes, int device, covo_flow** out) {
    if (!cfg || !packed_weights || !out) return fail(COVO_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->dim_head != 64) return fail(COVO_ERR_INVALID, "dim_head=%d unsupported (64 only)", cfg->dim_head);
    if (cfg->depth < 2 || cfg->depth % 2) return fail(COVO_ERR_INVALID, "depth=%d must be even", cfg->depth);
    if (cfg->dim % 128 || cfg->dim > 1024) return fail(COVO_ERR_INVALID, "dim=%d unsupported (128..1024, multiple of 128)", cfg->dim);
    if (cfg->n_streams != 1 && cfg->n_streams != 2) return fail(COVO_ERR_INVALID, "n_streams=%d", cfg->n_streams);
    if (cfg->conv_pos_kernel != 31) return fail(COVO_ERR_INVALID, "conv_pos_kernel=%d unsupported (31 only)", cfg->conv_pos_kernel);
    if (cfg->dim_x % 2) return fail(COVO_ERR_INVALID, "dim_x=%d must be even", cfg->dim_x);
    DeviceGuard g(device);
    covo_flow* h = new covo_flow();
    h->cfg = *cfg;
    int rc = check_device(device, &h->di);
    if (rc == COVO_OK) rc = init_kernel_attrs();
    if (rc == COVO_OK) rc = h->w.load(packed_weights, bytes);
    if (rc == COVO_OK) rc = flow_bind_weights(h);
    if (rc == COVO_OK && cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking) != cudaSuccess)
        rc = fail(COVO_ERR_CUDA, "cudaStreamCreate failed");
    if (rc != COVO_OK) {
        h->w.release();
        delete h;
        return rc;
    }
    h->use_graph = !env_flag("COVO_NO_GRAPH");
    // COVO_GEMM_MC=2: GEMMs run as cluster pairs sharing the weight tile by TMA multicast (gemm_tc_pair_kernel).  Measured
    // equal to independent CTAs at C3 (596.7 vs 596.0 ms) and slower at C2 (28.4 vs 27.1 ms): the L2 -> SM fabric is not
    // what binds these GEMMs, so the default stays 1.
    if (const char* v = getenv("COVO_GEMM_MC")) h->di.gemm_mc = atoi(v) == 2 ? 2 : 1;
    if (const char* v = getenv("COVO_GEMM_CG")) h->di.gemm_cg = atoi(v);
    if (const char* v = getenv("COVO_FLOW_PERSISTENT")) h->persistent_mode = atoi(v);
    if (const char* v = getenv("COVO_FLOW_PERSISTENT_ROWS")) h->persistent_max_rows = atoi(v);
    if (const char* v = getenv("COVO_FLOW_PERSISTENT_BN")) h->persistent_bn = atoi(v);
    h->naive_attn = env_flag("COVO_DEBUG_NAIVE_ATTN");
    if (const char* v = getenv("COVO_FLOW_SERPENTINE")) h->serpentine = atoi(v) != 0;
    *out = h;
    return COVO_OK;
}

int covo_flow_destroy(covo_flow* h) {
    if (!h) return COVO_OK;
    DeviceGuard g(h->di.device);
    cudaDeviceSynchronize();
    for (FlowPlan* p : h->plans) flow_free_plan(p);
    if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    h->w.release();
    delete h;
    return COVO_OK;
}

size_t covo_flow_workspace_bytes(const covo_flow* h, int B, int N, int n_eval_times) {
    if (!h || B < 1 || N < 1) return 0;
    FlowPlan p;
    p.B = B;
    p.N = N;
    p.BN = B * N;
    p.M = 2 * p.BN;
    p.n_t = n_eval_times < 1 ? 1 : (n_eval_times > FLOW_MAX_TIMES ? FLOW_MAX_TIMES : n_eval_times);
    p.ws = nullptr;
    return flow_layout(h, p);
}

int covo_flow_set_step_size(covo_flow* h, float step_size) {
    if (!h) return fail(COVO_ERR_INVALID, "null handle");
    if (!(step_size >= 0.f) || step_size > 1.f) return fail(COVO_ERR_INVALID, "step_size=%g out of range (0 = uniform grid, else (0, 1])", step_size);
    h->step_size = step_size;
    return COVO_OK;
}

int covo_flow_launches_per_sample(const covo_flow* h, int method, int n_steps, float cond_scale) {
    if (!h) return 0;
    (void)cond_scale;
    const int half = h->cfg.depth / 2;
    const int per_net = 2 + 7 * h->cfg.depth + half + 2;
    const int nfe = flow_num_times(method, n_steps);
    return 2 + 1 + nfe * (per_net + 1);      // per-call prologue (embedding gather, e_const), state -> input, the evaluations
}

int covo_flow_last_launches(const covo_flow* h) { return h ? h->last_launches : 0; }

int covo_flow_sample(covo_flow* h, const int64_t* ids, const float* cond, const float* y0, float* out, int B, int N,
                     int method, int n_steps, float cond_scale, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !ids || !cond || !y0 || !out) return fail(COVO_ERR_INVALID, "null argument");
    DeviceGuard g(h->di.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FlowPlan* p = nullptr;
    COVO_TRY(flow_get_plan(h, B, N, method, n_steps, cond_scale, 0, workspace, workspace_bytes, &p));
    const covo_flow_cfg& c = h->cfg;
    const size_t bn = static_cast<size_t>(B) * N;
    COVO_CK(cudaMemcpyAsync(p->ids, ids, bn * c.n_streams * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    COVO_CK(cudaMemcpyAsync(p->cond, cond, bn * c.dim_in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    COVO_CK(cudaMemcpyAsync(p->x_state, y0, bn * c.dim_x * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (p->exec && !prof().on) {
        COVO_CK(cudaGraphLaunch(p->exec, st));
    } else {
        int launches = 0;
        COVO_TRY(flow_enqueue_sample(h, *p, st, &launches));
        p->launches = launches;
    }
    h->last_launches = p->launches;
    COVO_CK(cudaMemcpyAsync(out, p->x_state, bn * c.dim_x * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return COVO_OK;
}

int covo_flow_velocity(covo_flow* h, const int64_t* ids, const float* cond, const float* x, float t, float* v, int B,
                       int N, float cond_scale, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !ids || !cond || !x || !v) return fail(COVO_ERR_INVALID, "null argument");
    DeviceGuard g(h->di.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FlowPlan* p = nullptr;
    COVO_TRY(flow_get_plan(h, B, N, COVO_ODE_EULER, 1, cond_scale, 1, workspace, workspace_bytes, &p));
    const covo_flow_cfg& c = h->cfg;
    const size_t bn = static_cast<size_t>(B) * N;
    p->times[0] = t;
    COVO_CK(cudaMemcpyAsync(p->ids, ids, bn * c.n_streams * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    COVO_CK(cudaMemcpyAsync(p->cond, cond, bn * c.dim_in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    COVO_CK(cudaMemcpyAsync(p->x_in, x, bn * c.dim_x * sizeof(float), cudaMemcpyDeviceToDevice, st));
    int launches = 0;
    COVO_TRY(flow_enqueue_velocity(h, *p, st, &launches));
    COVO_CK(cudaMemcpyAsync(v, p->v_out, bn * c.dim_x * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return COVO_OK;
}

// ====================================================================================== hifigan
int covo_hifigan_create(const covo_hifigan_cfg* cfg, const void* packed_weights, size_t bytes, int device,
                        covo_hifigan** out) {
    if (!cfg || !packed_weights || !out) return fail(COVO_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->num_upsamples < 1 || cfg->num_upsamples > 8) return fail(COVO_ERR_INVALID, "num_upsamples=%d", cfg->num_upsamples);
    if (cfg->num_kernels < 1 || cfg->num_kernels > 3) return fail(COVO_ERR_INVALID, "num_kernels=%d (1..3 supported)", cfg->num_kernels);
    if (cfg->num_dilations < 1 || cfg->num_dilations > 4) return fail(COVO_ERR_INVALID, "num_dilations=%d", cfg->num_dilations);
    if (cfg->resblock_type != 1 && cfg->resblock_type != 2) return fail(COVO_ERR_INVALID, "resblock_type=%d", cfg->resblock_type);
    for (int i = 0; i < cfg->num_upsamples; ++i) {
        const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
        if (u < 1 || k < u || (k + u - 1) / u > GEMM_MAX_TAPS) return fail(COVO_ERR_INVALID, "upsample %d: rate %d kernel %d unsupported", i, u, k);
    }
    for (int j = 0; j < cfg->num_kernels; ++j)
        if (cfg->resblock_kernel_sizes[j] > GEMM_MAX_TAPS || cfg->resblock_kernel_sizes[j] % 2 == 0)
            return fail(COVO_ERR_INVALID, "resblock kernel size %d unsupported (odd, <= %d)", cfg->resblock_kernel_sizes[j], GEMM_MAX_TAPS);
    DeviceGuard g(device);
    covo_hifigan* h = new covo_hifigan();
    h->cfg = *cfg;
    h->is_fp16 = cfg->h_format == COVO_H_FP16;
    h->allow_fused = !env_flag("COVO_HIFIGAN_NO_FUSED");
    h->mel_pad = pad64(cfg->num_mels);
    int rc = check_device(device, &h->di);
    if (rc == COVO_OK) rc = init_kernel_attrs();
    if (rc == COVO_OK) rc = h->w.load(packed_weights, bytes);
    if (rc != COVO_OK) {
        h->w.release();
        delete h;
        return rc;
    }
    if (const char* v = getenv("COVO_GEMM_CG")) h->di.gemm_cg = atoi(v);       // A/B switch, as for the flow handle
    rc = hifi_build_aligned_upconvs(h);
    if (rc != COVO_OK) {
        for (void* q : h->up_al) if (q) cudaFree(q);
        h->w.release();
        delete h;
        return rc;
    }
    *out = h;
    return COVO_OK;
}

int covo_hifigan_destroy(covo_hifigan* h) {
    if (!h) return COVO_OK;
    DeviceGuard g(h->di.device);
    cudaDeviceSynchronize();
    for (HifiPlan* p : h->plans) delete p;
    for (void* q : h->up_al) if (q) cudaFree(q);
    h->w.release();
    delete h;
    return COVO_OK;
}

size_t covo_hifigan_workspace_bytes(const covo_hifigan* h, int B, int T) {
    if (!h || B < 1 || T < 1) return 0;
    HifiPlan p;
    p.B = B;
    p.T = T;
    p.ws = nullptr;
    return hifi_layout(h, p);
}

int64_t covo_hifigan_out_len(const covo_hifigan* h, int T) { return h ? hifi_out_len(h->cfg, T) : 0; }

int covo_hifigan_launches_per_forward(const covo_hifigan* h) {
    if (!h) return 0;
    const covo_hifigan_cfg& c = h->cfg;
    const int per_rb = c.num_dilations * (c.resblock_type == 1 ? 2 : 1);
    const int layered = 2 + c.num_upsamples * (1 + c.num_kernels * per_rb + 1) + 1;
    // fused last stage: its convs, the stage mean and conv_post are one launch
    return hifi_fused_eligible(h) ? layered - (c.num_kernels * per_rb + 2) + 1 : layered;
}

int covo_hifigan_forward(covo_hifigan* h, const float* mel, void* wav, int B, int T, int out_dtype, void* workspace,
                         size_t workspace_bytes, void* stream) {
    if (!h || !mel || !wav) return fail(COVO_ERR_INVALID, "null argument");
    if (out_dtype < COVO_WAV_F32 || out_dtype > COVO_WAV_I16) return fail(COVO_ERR_INVALID, "out_dtype=%d", out_dtype);
    DeviceGuard g(h->di.device);
    HifiPlan* p = nullptr;
    COVO_TRY(hifi_get_plan(h, B, T, workspace, workspace_bytes, &p));
    return hifi_enqueue(h, *p, mel, wav, out_dtype, static_cast<cudaStream_t>(stream));
}

// ====================================================================================== text-to-semantic
int covo_t2s_create(const covo_t2s_cfg* cfg, const void* packed_weights, size_t bytes, int device, covo_t2s** out) {
    if (!cfg || !packed_weights || !out) return fail(COVO_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->dim_head != T2S_DH) return fail(COVO_ERR_INVALID, "dim_head=%d unsupported (64 only)", cfg->dim_head);
    if (cfg->target_depth < 1 || cfg->target_depth > T2S_MAX_DEPTH || cfg->source_depth < 0)
        return fail(COVO_ERR_INVALID, "target_depth=%d / source_depth=%d out of range", cfg->target_depth, cfg->source_depth);
    if (cfg->dim % 8 || cfg->target_transformer_dim % 64 || (cfg->heads * cfg->dim_head) % 64 || cfg->heads < 1 || (cfg->num_semantic_token_ids + 1) % 2 ||
        cfg->num_semantic_token_ids + 1 > T2S_THREADS)
        return fail(COVO_ERR_INVALID, "unsupported dims (dim=%d, target dim=%d, heads=%d, semantic ids=%d)", cfg->dim,
                    cfg->target_transformer_dim, cfg->heads, cfg->num_semantic_token_ids);
    if (cfg->weight_format != COVO_T2S_W_BF16 && cfg->weight_format != COVO_T2S_W_F32)
        return fail(COVO_ERR_INVALID, "weight_format=%d", cfg->weight_format);
    DeviceGuard g(device);
    covo_t2s* h = new covo_t2s();
    h->cfg = *cfg;
    int rc = check_device(device, &h->di);
    if (rc == COVO_OK) rc = h->w.load(packed_weights, bytes);
    if (rc == COVO_OK) rc = t2s_bind_weights(h);
    if (rc == COVO_OK) {
        int coop = 0;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop)
            rc = fail(COVO_ERR_CUDA, "device %d does not support cooperative launches", device);
    }
    if (rc != COVO_OK) {
        h->w.release();
        delete h;
        return rc;
    }
    *out = h;
    return COVO_OK;
}

int covo_t2s_destroy(covo_t2s* h) {
    if (!h) return COVO_OK;
    DeviceGuard g(h->di.device);
    cudaDeviceSynchronize();
    h->w.release();
    delete h;
    return COVO_OK;
}

size_t covo_t2s_workspace_bytes(const covo_t2s* h, int B, int S, int max_length) {
    if (!h || B < 1 || B > 8 || S < 1 || max_length < 1) return 0;
    return t2s_layout(h, t2s_pad_batch(B), S, max_length, nullptr, nullptr);
}

int covo_t2s_launches_per_generate(const covo_t2s* h) { return h ? t2s_source_launches(h) + 1 : 0; }

size_t covo_t2s_weight_bytes_per_step(const covo_t2s* h) {
    if (!h) return 0;
    const size_t Dt = h->cfg.target_transformer_dim, inner = h->inner;
    const size_t per_layer = 3 * inner * Dt + Dt * inner + inner * Dt + Dt * inner + 2 * static_cast<size_t>(h->ffi) * Dt +
                             Dt * static_cast<size_t>(h->ffi);
    const size_t esz = h->wdt == DT_F32 ? 4 : 2;
    return per_layer * h->cfg.target_depth * esz + static_cast<size_t>(h->n_logits) * h->demb * 4;
}

int covo_t2s_generate(covo_t2s* h, const int64_t* text_ids, const float* u, const int64_t* forced, int64_t* tokens,
                      int32_t* result, float* logits_out, float* enc_out, int B, int S, int max_length, float temperature,
                      int top_k, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !text_ids || !u || !tokens || !result || !workspace) return fail(COVO_ERR_INVALID, "null argument");
    if (B != 1 && B != 2 && B != 4 && B != 8)
        return fail(COVO_ERR_INVALID, "B=%d: the decode kernel takes 1, 2, 4 or 8 rows (the host mirror pads)", B);
    if (S < 1 || max_length < 1 || max_length > 8192) return fail(COVO_ERR_INVALID, "S=%d / max_length=%d out of range", S, max_length);
    if (top_k < 1 || top_k > h->n_logits) return fail(COVO_ERR_INVALID, "top_k=%d", top_k);
    DeviceGuard g(h->di.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const covo_t2s_cfg& c = h->cfg;
    T2SBuffers t;
    const size_t need = t2s_layout(h, B, S, max_length, workspace, &t);
    if (need > workspace_bytes) return fail(COVO_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, need);
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(COVO_ERR_INVALID, "workspace must be 256-byte aligned");
    COVO_CK(cudaMemcpyAsync(t.ids, text_ids, static_cast<size_t>(B) * S * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    COVO_CK(cudaMemsetAsync(static_cast<uint8_t*>(workspace) + t.zero_from, 0, t.zero_bytes, st));
    COVO_TRY(t2s_enqueue_source(h, t, B, S, max_length, st));
    if (enc_out)
        COVO_CK(cudaMemcpyAsync(enc_out, t.he, static_cast<size_t>(B) * S * c.dim * sizeof(float), cudaMemcpyDeviceToDevice, st));

    T2SDecArgs a;
    memset(&a, 0, sizeof(a));
    int trace_step = -1;
    const int H = c.heads, n_ctx = S + 1;
    const int n_ctx_r = round_up(n_ctx, 32), max_len_r = round_up(max_length, 32);
    const size_t ctx_per = static_cast<size_t>(B) * H * n_ctx_r * T2S_DH, cache_per = static_cast<size_t>(B) * H * max_len_r * T2S_DH;
    for (int L = 0; L < c.target_depth; ++L) {
        const T2SDecLayerW& d = h->dec[L];
        T2SLayerW& w = a.L[L];
        w.sa_gamma = d.sa_gamma.as<float>();
        w.sa_qkv = d.sa_qkv.ptr;
        w.sa_out = d.sa_out.ptr;
        w.ca_gamma = d.ca_gamma.as<float>();
        w.ca_q = d.ca_q.ptr;
        w.ca_out = d.ca_out.ptr;
        w.ff_gamma = d.ff_gamma.as<float>();
        w.ff1 = d.ff1_w.ptr;
        w.ff1_b = d.ff1_b.as<float>();
        w.ff2 = d.ff2_w.ptr;
        w.ff2_b = d.ff2_b.as<float>();
        w.ctx_k = t.ctx_k + L * ctx_per;
        w.ctx_v = t.ctx_v + L * ctx_per;
        w.kcache = t.kcache + L * cache_per;
        w.vcache = t.vcache + L * cache_per;
    }
    a.depth = c.target_depth;
    a.B = B;
    a.Dt = c.target_transformer_dim;
    a.inner = h->inner;
    a.H = H;
    a.ffi = h->ffi;
    a.ffi_pad = h->ffi_pad;
    a.n_out = h->n_out;
    a.demb = h->demb;
    a.n_logits = h->n_logits;
    a.n_ctx = n_ctx;
    a.max_len = max_length;
    a.n_ctx_r = n_ctx_r;
    a.max_len_r = max_len_r;
    a.topk = top_k;
    {
        const char* dm = getenv("COVO_T2S_DEBUG_SKIP");     // bit mask: 1 all work, 2 attention, 4 matrix products, 8 sampler, 16 no weight prefetch, 32 plain (not evict-first) weight loads
        a.dbg_mode = dm ? atoi(dm) : 0;
        const int skip_gemv = (a.dbg_mode & 4) ? 1 : 0;
        COVO_CK(cudaMemcpyToSymbolAsync(t2s_dbg_skip_gemv, &skip_gemv, sizeof(int), 0, cudaMemcpyHostToDevice, st));
        const char* ts = getenv("COVO_T2S_TRACE");
        trace_step = ts ? atoi(ts) : -1;
        COVO_CK(cudaMemcpyToSymbolAsync(t2s_trace_step, &trace_step, sizeof(int), 0, cudaMemcpyHostToDevice, st));
        const char* ag = getenv("COVO_T2S_ATTN_GROUPS");
        const int attn_groups = ag ? atoi(ag) : 0;
        COVO_CK(cudaMemcpyToSymbolAsync(t2s_dbg_attn_groups, &attn_groups, sizeof(int), 0, cudaMemcpyHostToDevice, st));
        const int plain = (a.dbg_mode & 32) ? 1 : 0;
        COVO_CK(cudaMemcpyToSymbolAsync(t2s_dbg_plain_loads, &plain, sizeof(int), 0, cudaMemcpyHostToDevice, st));
        const int no_pf = (a.dbg_mode & 16) ? 1 : 0;
        COVO_CK(cudaMemcpyToSymbolAsync(t2s_dbg_no_prefetch, &no_pf, sizeof(int), 0, cudaMemcpyHostToDevice, st));
    }
    a.temperature = temperature;
    a.ignore_eos = (flags & COVO_T2S_IGNORE_EOS) ? 1 : 0;
    a.eos_id = c.num_semantic_token_ids;
    a.emb = h->dec_emb.as<float>();
    a.start = h->dec_start.as<float>();
    a.final_gamma = h->dec_final.as<float>();
    a.rope = t.rope;
    a.ctx_mask = t.cmask;
    a.x = t.x;
    a.q = t.q;
    a.attn = t.attn;
    a.part = t.part;
    a.part_cnt = t.part_cnt;
    a.hbuf = t.hbuf;
    a.logits = t.logits;
    a.u = u;
    a.forced = reinterpret_cast<const long long*>(forced);
    a.tokens = reinterpret_cast<long long*>(tokens);
    a.logits_out = logits_out;
    a.result = t.result;
    a.eos_flags = t.eos_flags;
    a.barrier = t.barrier;
    const size_t smem = t2s_decode_smem(h, B);
    {
        ProfScope ps(PC_T2S_DECODE, 0.0, st);
        int rc = COVO_OK;
#define COVO_T2S_LAUNCH(WT_)                                                     \
    do {                                                                         \
        if (B == 1) rc = t2s_launch_decode<WT_, 1>(h, a, smem, st);              \
        else if (B == 2) rc = t2s_launch_decode<WT_, 2>(h, a, smem, st);         \
        else if (B == 4) rc = t2s_launch_decode<WT_, 4>(h, a, smem, st);         \
        else rc = t2s_launch_decode<WT_, 8>(h, a, smem, st);                     \
    } while (0)
        if (h->wdt == DT_F32) COVO_T2S_LAUNCH(float);
        else COVO_T2S_LAUNCH(__nv_bfloat16);
#undef COVO_T2S_LAUNCH
        COVO_TRY(rc);
    }
    COVO_CK(cudaMemcpyAsync(result, t.result, 4 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (trace_step >= 0) {      // debug only: synchronises
        long long tb[128];
        COVO_CK(cudaStreamSynchronize(st));
        COVO_CK(cudaMemcpyFromSymbol(tb, t2s_trace_buf, sizeof(tb)));
        static const char* names[21] = {"S1 norm", "S1 gemv", "S1 barrier", "S2 attn", "S2 barrier", "S3 load", "S3 gemv",
                                        "S3 barrier", "S4 norm+gemv", "S4 barrier", "S5 attn", "S5 barrier", "S6 load+gemv",
                                        "S6 barrier", "S7 norm", "S7 gemv", "S7 barrier", "S8 load", "S8 gemv", "S8 barrier", ""};
        for (int L = 0; L < c.target_depth; ++L) {
            fprintf(stderr, "[t2s trace] step %d layer %d (cycles):", trace_step, L);
            for (int i = 0; i < 20; ++i) fprintf(stderr, " %s=%lld", names[i], tb[2 + L * 24 + i + 1] - tb[2 + L * 24 + i]);
            fprintf(stderr, "\n");
        }
        for (int k = 0; k < 2; ++k) {
            fprintf(stderr, "[t2s trace] layer 0 %s attention, CTA 0 (cycles since stage start):", k ? "cross" : "self");
            for (int i = 1; i < 10; ++i) fprintf(stderr, " m%d=%lld", i, tb[106 + 10 * k + i] ? tb[106 + 10 * k + i] - tb[106 + 10 * k] : -1);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[t2s trace] S9 logits=%lld barrier=%lld S10 sample=%lld barrier=%lld\n", tb[101] - tb[100],
                tb[102] - tb[101], tb[103] - tb[102], tb[104] - tb[103]);
    }
    return COVO_OK;
}

// ====================================================================================== mel front-end
int covo_mel_create(const covo_mel_cfg* cfg, const float* window, const float* mel_basis, int device, covo_mel** out) {
    if (!cfg || !window || !mel_basis || !out) return fail(COVO_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->n_fft < 2 || cfg->n_fft > 4096 || cfg->win_size < 1 || cfg->win_size > cfg->n_fft || cfg->hop_size < 1 ||
        cfg->hop_size > cfg->n_fft || cfg->num_mels < 1)
        return fail(COVO_ERR_INVALID, "unsupported mel config (n_fft=%d, hop=%d, win=%d, mels=%d)", cfg->n_fft, cfg->hop_size,
                    cfg->win_size, cfg->num_mels);
    DeviceGuard g(device);
    covo_mel* h = new covo_mel();
    h->cfg = *cfg;
    h->n_freq = cfg->n_fft / 2 + 1;
    int rc = check_device(device, &h->di);
    if (rc == COVO_OK) {
        std::vector<float2> tw(cfg->n_fft);
        for (int j = 0; j < cfg->n_fft; ++j) {
            const double ang = 2.0 * 3.14159265358979323846 * j / cfg->n_fft;
            tw[j] = make_float2(static_cast<float>(cos(ang)), static_cast<float>(sin(ang)));
        }
        const size_t nb = static_cast<size_t>(cfg->num_mels) * h->n_freq * sizeof(float);
        if (cudaMalloc(&h->window, cfg->win_size * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&h->twiddle, cfg->n_fft * sizeof(float2)) != cudaSuccess || cudaMalloc(&h->basis, nb) != cudaSuccess ||
            cudaMemcpy(h->window, window, cfg->win_size * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->twiddle, tw.data(), cfg->n_fft * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->basis, mel_basis, nb, cudaMemcpyHostToDevice) != cudaSuccess)
            rc = fail(COVO_ERR_CUDA, "mel front-end: device allocation / copy failed");
    }
    if (rc != COVO_OK) {
        covo_mel_destroy(h);
        return rc;
    }
    *out = h;
    return COVO_OK;
}

int covo_mel_destroy(covo_mel* h) {
    if (!h) return COVO_OK;
    DeviceGuard g(h->di.device);
    cudaDeviceSynchronize();
    if (h->window) cudaFree(h->window);
    if (h->twiddle) cudaFree(h->twiddle);
    if (h->basis) cudaFree(h->basis);
    delete h;
    return COVO_OK;
}

int covo_mel_frames(const covo_mel* h, int L) {
    if (!h) return 0;
    const int pad = (h->cfg.n_fft - h->cfg.hop_size) / 2;
    if (L < pad + 1 || L + 2 * pad < h->cfg.n_fft) return 0;
    return (L + 2 * pad - h->cfg.n_fft) / h->cfg.hop_size + 1;
}

int covo_mel_forward(covo_mel* h, const float* wav, float* mel, int B, int L, void* stream) {
    if (!h || !wav || !mel) return fail(COVO_ERR_INVALID, "null argument");
    const int T = covo_mel_frames(h, L);
    if (B < 1 || B > 65535 || T < 1) return fail(COVO_ERR_INVALID, "mel front-end: B=%d, L=%d gives no frame (reflect padding needs L > %d)", B, L,
                                    (h->cfg.n_fft - h->cfg.hop_size) / 2);
    DeviceGuard g(h->di.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MelArgs a;
    a.wav = wav;
    a.mel = mel;
    a.window = h->window;
    a.twiddle = h->twiddle;
    a.basis = h->basis;
    a.L = L;
    a.T = T;
    a.n_fft = h->cfg.n_fft;
    a.hop = h->cfg.hop_size;
    a.win = h->cfg.win_size;
    a.n_freq = h->n_freq;
    a.n_mels = h->cfg.num_mels;
    a.pad = (h->cfg.n_fft - h->cfg.hop_size) / 2;
    const size_t smem = sizeof(float) * (3 * static_cast<size_t>(a.n_fft) + a.n_freq);
    if (smem > 48 * 1024)
        COVO_CK(cudaFuncSetAttribute(mel_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    ProfScope ps(PC_PROLOGUE, 0.0, st);
    mel_frontend_kernel<<<dim3(T, B), 256, smem, st>>>(a);
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

// ====================================================================================== SM budgets
namespace {
int apply_sm_limit(DeviceInfo* di, int n_sms) {
    cudaDeviceProp p;
    COVO_CK(cudaGetDeviceProperties(&p, di->device));
    di->num_sms = (n_sms <= 0 || n_sms > p.multiProcessorCount) ? p.multiProcessorCount : n_sms;
    return COVO_OK;
}
}  // namespace
int covo_flow_set_sm_limit(covo_flow* h, int n_sms) { return h ? apply_sm_limit(&h->di, n_sms) : fail(COVO_ERR_INVALID, "null handle"); }
int covo_hifigan_set_sm_limit(covo_hifigan* h, int n_sms) { return h ? apply_sm_limit(&h->di, n_sms) : fail(COVO_ERR_INVALID, "null handle"); }
int covo_t2s_set_sm_limit(covo_t2s* h, int n_sms) { return h ? apply_sm_limit(&h->di, n_sms) : fail(COVO_ERR_INVALID, "null handle"); }

// ====================================================================================== profiler
int covo_prof_begin(void) {
    Profiler& p = prof();
    p.recs.clear();
    p.on = true;
    return COVO_OK;
}

int covo_prof_end(double* ms_per_class, double* flops_per_class, int* launches_per_class, int n_classes) {
    Profiler& p = prof();
    p.on = false;
    COVO_CK(cudaDeviceSynchronize());
    for (int i = 0; i < n_classes; ++i) {
        ms_per_class[i] = 0.0;
        flops_per_class[i] = 0.0;
        launches_per_class[i] = 0;
    }
    for (ProfRec& r : p.recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (r.cat < n_classes) {
            ms_per_class[r.cat] += ms;
            flops_per_class[r.cat] += r.flops;
            launches_per_class[r.cat] += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    p.recs.clear();
    return COVO_OK;
}

// ====================================================================================== test hooks
static int dbg_device(DeviceInfo* di) {
    static bool ready = false;
    static DeviceInfo cached;
    if (!ready) {
        int dev = 0;
        COVO_CK(cudaGetDevice(&dev));
        COVO_TRY(check_device(dev, &cached));
        COVO_TRY(init_kernel_attrs());
        ready = true;
    }
    *di = cached;
    return COVO_OK;
}

int covo_dbg_gemm(const void* A_bf16, const void* W_bf16, const float* bias, const float* residual, float* out_f32,
                  void* out_bf16, int M, int N, int K, int act_h, int force_bn, void* stream) {
    DeviceInfo di;
    COVO_TRY(dbg_device(&di));
    if (N % 64 || K % 64) return fail(COVO_ERR_INVALID, "dbg_gemm needs N, K multiples of 64");
    if (const char* v = getenv("COVO_GEMM_MC")) di.gemm_mc = atoi(v) == 2 ? 2 : 1;
    if (const char* v = getenv("COVO_GEMM_CG")) di.gemm_cg = atoi(v);
    GemmOp op;
    gemm_defaults(op.args);
    COVO_TRY(build_gemm(op, di, a2d(A_bf16, K, M), M, 1, W_bf16, N, 1, 0, force_bn));
    COVO_TRY(gemm_set_outputs(op, out_f32, residual, out_bf16, N, M, 1, N, 0, 0));
    op.args.bias = bias;
    op.args.act_h = act_h;
    if (getenv("COVO_GEMM_TRACE") == nullptr) return launch_gemm(op, static_cast<cudaStream_t>(stream));
    // debug: CTA 0's clock64 timeline of this launch (gemm_sm100.cuh, g_gemm_trace); synchronises
    long long* d = nullptr;
    long long hrec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int zero = 0;
    COVO_CK(cudaMalloc(&d, 8 * 200 * sizeof(long long)));
    COVO_CK(cudaMemset(d, 0, 8 * 200 * sizeof(long long)));
    COVO_CK(cudaMemcpyToSymbol(g_gemm_trace, &d, sizeof(d)));
    COVO_CK(cudaMemcpyToSymbol(g_gemm_trace_n, &zero, sizeof(zero)));
    int rc = launch_gemm(op, static_cast<cudaStream_t>(stream));
    cudaDeviceSynchronize();
    cudaMemcpy(hrec, d, sizeof(hrec), cudaMemcpyDeviceToHost);
    long long* nul = nullptr;
    cudaMemcpyToSymbol(g_gemm_trace, &nul, sizeof(nul));
    cudaFree(d);
    const int m_tiles = ceil_div(M, GEMM_BM);
    const int tiles = (op.cg == 2 || op.mc == 2) ? ceil_div(m_tiles, 2) * (N / op.bn) : m_tiles * (N / op.bn);
    const int per_cta = ceil_div(tiles, (op.cg == 2 || op.mc == 2) ? op.grid / 2 : op.grid);
    fprintf(stderr, "[gemm trace] M=%d N=%d K=%d bn=%d cg=%d, %d tiles on CTA 0: init %lld | first stage +%lld | last MMA issued +%lld "
            "(%.0f cycles per tile) | last accumulator complete +%lld | its stores issued +%lld | complete +%lld | MMA thread waited "
            "%lld cycles for accumulator buffers\n", M, N, K, op.bn, op.cg, per_cta, hrec[1] - hrec[0], hrec[2] - hrec[1], hrec[3] - hrec[2],
            static_cast<double>(hrec[3] - hrec[2]) / per_cta, hrec[4] - hrec[3], hrec[5] - hrec[4], hrec[6] - hrec[5], hrec[7]);
    return rc;
}

int covo_dbg_attention(const void* qkv_bf16, void* out_bf16, int Bt, int N, int heads, int impl, void* stream) {
    DeviceInfo di;
    COVO_TRY(dbg_device(&di));
    const int inner = heads * 64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (impl == 1) {
        dim3 grid(ceil_div(N, 8), heads, Bt);
        naive_attention_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv_bf16),
                                                     static_cast<__nv_bfloat16*>(out_bf16), N, heads, inner, 0.125f);
    } else {
        AttnArgs a;
        COVO_TRY(attn_build_args(a, qkv_bf16, out_bf16, Bt, N, heads));
        COVO_TRY(launch_attention_kernel(a, di.num_sms, st));
    }
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

}  // extern "C"
