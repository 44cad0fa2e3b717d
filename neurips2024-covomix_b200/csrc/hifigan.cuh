// HiFi-GAN generator: host-side plan for Generator.forward (hifi-gan/models.py:100-116).
// Activations are time-major [B, T, C_pad] (channels contiguous, padded to a multiple of 64 with zeros),
// 16-bit (bf16 or fp16) as MMA operands plus an fp32 residual stream.  Every conv is one launch of the
// tcgen05 implicit-GEMM kernel (gemm_sm100.cuh) with LeakyReLU / bias / residual fused in its epilogue.
#pragma once
#include "common.cuh"
#include "hifigan_fused.cuh"

namespace covo {

struct HifiStage {
    int c_in_pad, c_out_pad, c_out;
    int t_in, t_out;
    int stride, ksize, pad, jtaps;
    float* x = nullptr;            // fp32 [B, t_out, c_out_pad] : ConvTranspose1d output (residual start)
    uint16_t* a0 = nullptr;        // 16-bit lrelu(x)
    float* xr[4] = {nullptr, nullptr, nullptr, nullptr};      // per-resblock fp32 stream
    uint16_t* ar[4] = {nullptr, nullptr, nullptr, nullptr};   // lrelu(xr)
    uint16_t* hr[4] = {nullptr, nullptr, nullptr, nullptr};   // lrelu(conv1 out)
    uint16_t* out_act = nullptr;   // lrelu(mean of resblocks): input of the next stage / conv_post
    GemmOp up;
    std::vector<GemmOp> convs;     // in launch order
};

struct HifiPlan {
    int B = 0, T = 0;
    void* ws = nullptr;
    uint16_t* mel_tc = nullptr;    // [B, T, mel_pad]
    uint16_t* pre_act = nullptr;   // lrelu(conv_pre) [B, T, c0_pad]
    GemmOp pre;
    std::vector<HifiStage> stages;
    int launches = 0;
    bool fused_last = false;       // last stage (resblocks + mean + conv_post + tanh) runs as hifigan_fused_last_stage_kernel
    HifiFusedArgs fused;
    double fused_flops = 0.0;
};

}  // namespace covo

struct covo_hifigan {
    covo_hifigan_cfg cfg;
    covo::DeviceInfo di;
    covo::Weights w;
    int mel_pad = 0;
    int is_fp16 = 0;
    bool allow_fused = true;       // COVO_HIFIGAN_NO_FUSED=1 keeps the layer-by-layer path for the last stage
    // ConvTranspose1d layers whose output length is stride * input length (k - 2 pad == stride): "aligned polyphase" weights
    // built once at creation (hifi_build_aligned_upconvs); null -> polyphase weights of the blob + per-element scatter epilogue
    void* up_al[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int up_al_taps[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int up_al_m0[8] = {0, 0, 0, 0, 0, 0, 0, 0};         // first tap index m (-1 when pad > 0, else 0)
    std::vector<covo::HifiPlan*> plans;
};

namespace covo {

inline int hifi_chan(const covo_hifigan_cfg& c, int stage /* -1: conv_pre output */) {
    return c.upsample_initial_channel >> (stage + 1);
}
inline int pad64(int c) { return round_up(c, 64); }

inline int64_t hifi_out_len(const covo_hifigan_cfg& c, int T) {
    int64_t L = T;
    for (int i = 0; i < c.num_upsamples; ++i) {
        const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i], p = (k - u) / 2;
        L = (L - 1) * u - 2 * p + k;
    }
    return L;
}

// Buffers of one stage are dead once the next stage's ConvTranspose1d has consumed `out_act`, so all stages share
// one scratch region (sized for the largest stage) and the stage outputs ping-pong between two regions.
inline size_t hifi_layout(const covo_hifigan* h, HifiPlan& p) {
    const covo_hifigan_cfg& c = h->cfg;
    Arena a(p.ws, static_cast<size_t>(-1));
    const size_t B = p.B;
    p.mel_tc = a.take<uint16_t>(B * p.T * h->mel_pad);
    p.pre_act = a.take<uint16_t>(B * p.T * pad64(hifi_chan(c, -1)));
    p.stages.assign(c.num_upsamples, HifiStage());
    int t = p.T;
    size_t max_n = 0;
    for (int i = 0; i < c.num_upsamples; ++i) {
        HifiStage& s = p.stages[i];
        s.stride = c.upsample_rates[i];
        s.ksize = c.upsample_kernel_sizes[i];
        s.pad = (s.ksize - s.stride) / 2;
        s.jtaps = ceil_div(s.ksize, s.stride);
        s.c_in_pad = pad64(hifi_chan(c, i - 1));
        s.c_out = hifi_chan(c, i);
        s.c_out_pad = pad64(s.c_out);
        s.t_in = t;
        s.t_out = (t - 1) * s.stride - 2 * s.pad + s.ksize;
        t = s.t_out;
        const size_t n = B * s.t_out * s.c_out_pad;
        if (n > max_n) max_n = n;
    }
    uint16_t* out_pp[2] = {a.take<uint16_t>(max_n), a.take<uint16_t>(max_n)};
    a.off = align_up(a.off, 256);
    const size_t scratch0 = a.off;
    size_t scratch_end = scratch0;
    for (int i = 0; i < c.num_upsamples; ++i) {
        HifiStage& s = p.stages[i];
        const size_t n = B * s.t_out * s.c_out_pad;
        a.off = scratch0;
        s.x = a.take<float>(n);
        s.a0 = a.take<uint16_t>(n);
        for (int j = 0; j < c.num_kernels; ++j) {
            s.xr[j] = a.take<float>(n);
            s.ar[j] = a.take<uint16_t>(n);
            s.hr[j] = a.take<uint16_t>(n);
        }
        s.out_act = out_pp[i & 1];
        if (a.off > scratch_end) scratch_end = a.off;
    }
    return align_up(scratch_end, 256);
}

inline ASource a3d(const void* ptr, int C, int T, int B) {
    ASource a;
    a.ptr = ptr;
    a.K = C;
    a.rows = T;
    a.Z = B;
    a.row_stride = C;
    a.z_stride = static_cast<long long>(C) * T;
    return a;
}

// Aligned polyphase form of ConvTranspose1d(stride s, kernel K, padding p) with K - 2p == s (t_out == s * t_in):
//     out[s q + rr, co] = sum_m sum_ci x[q - m, ci] W[ci, co, s m + rr + p],      m in [m0, m0 + taps)
// so row q of the GEMM holds the s output samples s q .. s q + s - 1 and the output [B][t_out][C] is simply the GEMM's
// [B][t_in][s * C] box -- ordinary TMA sub-tile stores, no scatter.  (The blob's polyphase form indexes phases by
// r = (o + p) mod s, which puts sample o = s q + r - p of row q into output row q or q - 1 depending on the phase.)
// Built from the blob's matrix pw[(r * Cop + co)][j * Cip + ci] = W[ci, co, r + s j]:  r = (rr + p) mod s, j = m + (rr + p) / s.
__global__ void upconv_align_weights_kernel(const uint16_t* __restrict__ pw, uint16_t* __restrict__ al, int s, int pad, int J,
                                            int taps, int m0, int cop, int cip) {
    const long long total = static_cast<long long>(s) * cop * taps * cip;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % cip);
        const int mt = static_cast<int>((i / cip) % taps);
        const int co = static_cast<int>((i / (static_cast<long long>(cip) * taps)) % cop);
        const int rr = static_cast<int>(i / (static_cast<long long>(cip) * taps * cop));
        const int r = (rr + pad) % s, j = m0 + mt + (rr + pad) / s;
        al[i] = (j >= 0 && j < J) ? pw[(static_cast<long long>(r) * cop + co) * (static_cast<long long>(J) * cip) + static_cast<long long>(j) * cip + ci]
                                  : static_cast<uint16_t>(0);
    }
}

inline bool env_flag_hifi(const char* name) {
    const char* v = getenv(name);
    return v && v[0] && v[0] != '0';
}

inline int hifi_build_aligned_upconvs(covo_hifigan* h) {
    const covo_hifigan_cfg& c = h->cfg;
    if (env_flag_hifi("COVO_HIFIGAN_SCATTER")) return COVO_OK;           // A/B switch: keep the scatter epilogue everywhere
    const int hdt = h->is_fp16 ? DT_F16 : DT_BF16;
    for (int i = 0; i < c.num_upsamples && i < 8; ++i) {
        const int s = c.upsample_rates[i], k = c.upsample_kernel_sizes[i], pad = (k - s) / 2;
        if (k - 2 * pad != s) continue;                                  // t_out != s * t_in (e.g. rate 5, kernel 8): scatter path
        const int J = ceil_div(k, s), m0 = pad > 0 ? -1 : 0;
        int m_hi = 0;                                                    // largest m with s m + rr + pad < k for some rr
        for (int rr = 0; rr < s; ++rr) m_hi = std::max(m_hi, (k - 1 - rr - pad) / s);
        const int taps = m_hi - m0 + 1;
        if (taps > GEMM_MAX_TAPS) continue;
        const int cip = pad64(hifi_chan(c, i - 1)), cop = pad64(hifi_chan(c, i));
        Tensor tw;
        COVO_TRY(h->w.get("ups." + std::to_string(i) + ".w", hdt, &tw));
        if (tw.shape[0] != static_cast<int64_t>(s) * cop || tw.shape[1] != static_cast<int64_t>(J) * cip)
            return fail(COVO_ERR_WEIGHTS, "ups.%d.w has unexpected shape", i);
        const size_t n = static_cast<size_t>(s) * cop * taps * cip;
        COVO_CK(cudaMalloc(&h->up_al[i], n * 2));
        upconv_align_weights_kernel<<<static_cast<int>(std::min<size_t>((n + 255) / 256, 4096)), 256>>>(
            tw.as<uint16_t>(), static_cast<uint16_t*>(h->up_al[i]), s, pad, J, taps, m0, cop, cip);
        COVO_CK(cudaGetLastError());
        h->up_al_taps[i] = taps;
        h->up_al_m0[i] = m0;
    }
    COVO_CK(cudaDeviceSynchronize());
    return COVO_OK;
}

// Conv1d(C->C_out, k, dilation d, "same" padding) as taps over time-major activations.
inline int hifi_conv_op(const covo_hifigan* h, GemmOp& op, const void* act, int c_in_pad, int T, int B, const Tensor& w,
                        const Tensor& bias, int c_out_pad, int k, int dil) {
    gemm_defaults(op.args);
    COVO_TRY(build_gemm(op, h->di, a3d(act, c_in_pad, T, B), T, B, w.ptr, c_out_pad, k, h->is_fp16));
    const int pad = (k * dil - dil) / 2;                       // get_padding, hifi-gan/utils.py:34-35
    for (int j = 0; j < k; ++j) {
        op.args.tap_row[j] = j * dil - pad;
        op.args.tap_z[j] = 0;
    }
    op.args.bias = bias.as<float>();
    op.cat = PC_GEMM_VOC;
    return COVO_OK;
}
inline int hifi_conv_out(const covo_hifigan* h, GemmOp& op, float* out_f32, const float* residual, void* out_h, int T, int B,
                         int c_out_pad) {
    return gemm_set_outputs(op, out_f32, residual, out_h, c_out_pad, T, B, c_out_pad, static_cast<long long>(T) * c_out_pad,
                            h->is_fp16);
}
inline double conv_flops(int B, int T, int cin, int cout, int k) { return 2.0 * B * T * static_cast<double>(cin) * cout * k; }

// The fused last-stage kernel covers ResBlock1 stages of <= 32 channels whose receptive field fits its tile
// (config_covomix.json: 31 channels, k = 3/7/11, dilations 1/3/5: halo 60, tap reach 25).
inline bool hifi_fused_eligible(const covo_hifigan* h) {
    const covo_hifigan_cfg& c = h->cfg;
    if (!h->allow_fused || c.resblock_type != 1 || c.num_kernels > 3 || c.num_dilations > 4) return false;
    if (hifi_chan(c, c.num_upsamples - 1) > HF_C) return false;
    for (int j = 0; j < c.num_kernels; ++j) {
        const int k = c.resblock_kernel_sizes[j];
        if (k > HF_MAXK || k % 2 == 0) return false;
        int halo = 0;
        for (int m = 0; m < c.num_dilations; ++m) {
            const int d = c.resblock_dilations[j][m];
            if ((k - 1) / 2 * d > HF_MARGIN) return false;
            halo += (k - 1) / 2 * d + (k - 1) / 2;
        }
        if (2 * (halo + 3) + HF_TP > HF_R) return false;
    }
    return true;
}

inline int hifi_build_ops(covo_hifigan* h, HifiPlan& p) {
    const covo_hifigan_cfg& c = h->cfg;
    const Weights& w = h->w;
    const uint32_t hdt = h->is_fp16 ? DT_F16 : DT_BF16;
    Tensor tw, tb;
    COVO_TRY(w.get("conv_pre.w", hdt, &tw));
    COVO_TRY(w.get("conv_pre.b", DT_F32, &tb));
    const int c0p = pad64(hifi_chan(c, -1));
    COVO_TRY(hifi_conv_op(h, p.pre, p.mel_tc, h->mel_pad, p.T, p.B, tw, tb, c0p, 7, 1));
    COVO_TRY(hifi_conv_out(h, p.pre, nullptr, nullptr, p.pre_act, p.T, p.B, c0p));    // x = lrelu(conv_pre(mel)) (models.py:101-103)
    p.pre.args.act_h = ACT_LRELU;
    p.pre.args.slope = 0.1f;
    p.pre.flops = conv_flops(p.B, p.T, c.num_mels, hifi_chan(c, -1), 7);
    p.launches = 2;                                             // mel transpose + conv_pre

    for (int i = 0; i < c.num_upsamples; ++i) {
        HifiStage& s = p.stages[i];
        const void* in_act = (i == 0) ? p.pre_act : p.stages[i - 1].out_act;
        // ---- ConvTranspose1d, polyphase: output sample o = stride*q + r - pad, phase r in [0, stride);
        //      D[q, r*C + co] = sum_j x[q - j] W[:, co, r + stride*j]   (weights packed [stride*C_pad, jtaps*Cin_pad])
        {
            const std::string nm = "ups." + std::to_string(i);
            COVO_TRY(w.get(nm + ".w", hdt, &tw));
            COVO_TRY(w.get(nm + ".b", DT_F32, &tb));
            GemmOp& op = s.up;
            gemm_defaults(op.args);
            COVO_TRY(build_gemm(op, h->di, a3d(in_act, s.c_in_pad, s.t_in, p.B), s.t_in + s.jtaps - 1, p.B, tw.ptr,
                                s.stride * s.c_out_pad, s.jtaps, h->is_fp16));
            for (int j = 0; j < s.jtaps; ++j) {
                op.args.tap_row[j] = -j;
                op.args.tap_z[j] = 0;
            }
            op.args.scatter = 1;
            op.args.n_valid = s.stride * s.c_out_pad;
            op.args.out_zs = static_cast<long long>(s.t_out) * s.c_out_pad;
            op.args.out_rs = static_cast<long long>(s.stride) * s.c_out_pad;
            op.args.out_off = -static_cast<long long>(s.pad) * s.c_out_pad;
            op.args.up_s = s.stride;
            op.args.up_p = s.pad;
            op.args.phase_w = s.c_out_pad;
            op.args.t_out = s.t_out;
            op.args.bias = tb.as<float>();
            op.args.out_f32 = s.x;
            op.args.out_h = s.a0;
            op.args.act_h = ACT_LRELU;
            op.args.slope = 0.1f;
            op.args.h_is_fp16 = h->is_fp16;
            op.flops = conv_flops(p.B, s.t_in, hifi_chan(c, i - 1), s.c_out, s.ksize);
            op.cat = PC_GEMM_VOC;
            ++p.launches;
        }
        const bool fused_next = i + 1 == c.num_upsamples && hifi_fused_eligible(h);
        if (fused_next) {
            // the fused kernel reads only the fp32 stream's first HF_C channels: the ConvTranspose1d need not write the
            // 16-bit copy nor the padding channels
            s.up.args.out_h = nullptr;
            s.up.args.scatter_c_valid = HF_C;
        }
        // t_out == stride * t_in: the aligned polyphase form (hifi_build_aligned_upconvs) -- an ordinary conv GEMM whose row q
        // holds output samples stride*q .. stride*q + stride - 1, stored by the TMA sub-tile path (the per-element scatter
        // epilogue ran these launches at ~1 TB/s of DRAM traffic)
        if (h->up_al[i] != nullptr && s.t_out == s.stride * s.t_in) {
            GemmOp& op = s.up;
            const float* bias = op.args.bias;
            void* out_h = op.args.out_h;
            const int c_valid = op.args.scatter_c_valid;
            const double flops = op.flops;
            gemm_defaults(op.args);
            COVO_TRY(build_gemm(op, h->di, a3d(in_act, s.c_in_pad, s.t_in, p.B), s.t_in, p.B, h->up_al[i], s.stride * s.c_out_pad,
                                h->up_al_taps[i], h->is_fp16));
            for (int mt = 0; mt < h->up_al_taps[i]; ++mt) {
                op.args.tap_row[mt] = -(h->up_al_m0[i] + mt);
                op.args.tap_z[mt] = 0;
            }
            COVO_TRY(gemm_set_outputs(op, s.x, nullptr, out_h, s.stride * s.c_out_pad, s.t_in, p.B,
                                      static_cast<long long>(s.stride) * s.c_out_pad, static_cast<long long>(s.t_out) * s.c_out_pad,
                                      h->is_fp16));
            op.args.bias = bias;
            op.args.act_h = ACT_LRELU;
            op.args.slope = 0.1f;
            op.args.phase_w = s.c_out_pad;
            op.args.scatter_c_valid = c_valid;
            op.flops = flops;
            op.cat = PC_GEMM_VOC;
        }
        // ---- narrow last stage: one fused kernel instead of 6 * num_kernels GEMM launches + mean + conv_post
        s.convs.clear();
        if (fused_next) {
            HifiFusedArgs& f = p.fused;
            memset(&f, 0, sizeof(f));
            f.x = s.x;
            f.nk = c.num_kernels;
            f.nd = c.num_dilations;
            f.T = s.t_out;
            f.ldx = s.c_out_pad;
            f.ld_post = s.c_out_pad;
            f.H = 0;
            f.inv_nk = 1.0f / static_cast<float>(c.num_kernels);
            f.slope_res = 0.1f;
            f.slope_post = 0.01f;                       // F.leaky_relu default before conv_post (models.py:112)
            p.fused_flops = conv_flops(p.B, s.t_out, s.c_out, 1, 7);
            for (int j = 0; j < c.num_kernels; ++j) {
                const int k = c.resblock_kernel_sizes[j];
                const int r = i * c.num_kernels + j;
                f.ksize[j] = k;
                f.halo[j] = 0;
                for (int m = 0; m < c.num_dilations; ++m) {
                    const int d = c.resblock_dilations[j][m];
                    f.dil[j][m] = d;
                    f.halo[j] += (k - 1) / 2 * d + (k - 1) / 2;
                    Tensor t1, t2, u1, u2;
                    COVO_TRY(w.get("rb." + std::to_string(r) + ".c1." + std::to_string(m) + ".w", hdt, &t1));
                    COVO_TRY(w.get("rb." + std::to_string(r) + ".c1." + std::to_string(m) + ".b", DT_F32, &u1));
                    COVO_TRY(w.get("rb." + std::to_string(r) + ".c2." + std::to_string(m) + ".w", hdt, &t2));
                    COVO_TRY(w.get("rb." + std::to_string(r) + ".c2." + std::to_string(m) + ".b", DT_F32, &u2));
                    f.w1[j][m] = t1.ptr;
                    f.w2[j][m] = t2.ptr;
                    f.b1[j][m] = u1.as<float>();
                    f.b2[j][m] = u2.as<float>();
                    p.fused_flops += 2.0 * conv_flops(p.B, s.t_out, s.c_out, s.c_out, k);
                }
                if (f.halo[j] > f.H) f.H = f.halo[j];
            }
            Tensor pw, pb;
            COVO_TRY(w.get("conv_post.w", DT_F32, &pw));
            COVO_TRY(w.get("conv_post.b", DT_F32, &pb));
            f.wpost = pw.as<float>();
            f.bpost = pb.as<float>();
            p.fused_last = true;
            p.launches += 1;
            continue;
        }
        // ---- resblocks
        for (int j = 0; j < c.num_kernels; ++j) {
            const int k = c.resblock_kernel_sizes[j];
            const int r = i * c.num_kernels + j;
            for (int m = 0; m < c.num_dilations; ++m) {
                const int d = c.resblock_dilations[j][m];
                const void* in1 = (m == 0) ? s.a0 : s.ar[j];
                const float* res = (m == 0) ? s.x : s.xr[j];
                if (c.resblock_type == 1) {
                    // xt = c2(lrelu(c1(lrelu(x)))); x = xt + x     (models.py:35-42)
                    const std::string n1 = "rb." + std::to_string(r) + ".c1." + std::to_string(m);
                    const std::string n2 = "rb." + std::to_string(r) + ".c2." + std::to_string(m);
                    COVO_TRY(w.get(n1 + ".w", hdt, &tw));
                    COVO_TRY(w.get(n1 + ".b", DT_F32, &tb));
                    GemmOp o1;
                    COVO_TRY(hifi_conv_op(h, o1, in1, s.c_out_pad, s.t_out, p.B, tw, tb, s.c_out_pad, k, d));
                    COVO_TRY(hifi_conv_out(h, o1, nullptr, nullptr, s.hr[j], s.t_out, p.B, s.c_out_pad));
                    o1.args.act_h = ACT_LRELU;
                    o1.flops = conv_flops(p.B, s.t_out, s.c_out, s.c_out, k);
                    s.convs.push_back(o1);
                    COVO_TRY(w.get(n2 + ".w", hdt, &tw));
                    COVO_TRY(w.get(n2 + ".b", DT_F32, &tb));
                    GemmOp o2;
                    COVO_TRY(hifi_conv_op(h, o2, s.hr[j], s.c_out_pad, s.t_out, p.B, tw, tb, s.c_out_pad, k, 1));
                    COVO_TRY(hifi_conv_out(h, o2, s.xr[j], res, (m + 1 < c.num_dilations) ? s.ar[j] : nullptr, s.t_out, p.B,
                                           s.c_out_pad));
                    o2.args.act_h = ACT_LRELU;
                    o2.flops = conv_flops(p.B, s.t_out, s.c_out, s.c_out, k);
                    s.convs.push_back(o2);
                } else {
                    // xt = c(lrelu(x)); x = xt + x                   (models.py:63-68)
                    const std::string n1 = "rb." + std::to_string(r) + ".c." + std::to_string(m);
                    COVO_TRY(w.get(n1 + ".w", hdt, &tw));
                    COVO_TRY(w.get(n1 + ".b", DT_F32, &tb));
                    GemmOp o1;
                    COVO_TRY(hifi_conv_op(h, o1, in1, s.c_out_pad, s.t_out, p.B, tw, tb, s.c_out_pad, k, d));
                    COVO_TRY(hifi_conv_out(h, o1, s.xr[j], res, (m + 1 < c.num_dilations) ? s.ar[j] : nullptr, s.t_out, p.B,
                                           s.c_out_pad));
                    o1.args.act_h = ACT_LRELU;
                    o1.flops = conv_flops(p.B, s.t_out, s.c_out, s.c_out, k);
                    s.convs.push_back(o1);
                }
            }
        }
        p.launches += static_cast<int>(s.convs.size()) + 1;     // + stage mean
    }
    if (!p.fused_last) p.launches += 1;                         // conv_post
    return COVO_OK;
}

inline int hifi_get_plan(covo_hifigan* h, int B, int T, void* ws, size_t ws_bytes, HifiPlan** out) {
    if (B < 1 || T < 1) return fail(COVO_ERR_INVALID, "B=%d T=%d must be positive", B, T);
    for (HifiPlan* q : h->plans)
        if (q->B == B && q->T == T && q->ws == ws) {
            *out = q;
            return COVO_OK;
        }
    HifiPlan* p = new HifiPlan();
    p->B = B;
    p->T = T;
    p->ws = ws;
    const size_t need = hifi_layout(h, *p);
    if (ws == nullptr || need > ws_bytes) {
        delete p;
        return fail(COVO_ERR_INVALID, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    }
    int rc = hifi_build_ops(h, *p);
    if (rc != COVO_OK) {
        delete p;
        return rc;
    }
    // padded channels of the mel operand must be exact zeros: mel_to_tc_kernel rewrites them on every call
    if (h->plans.size() >= 8) {
        delete h->plans.front();
        h->plans.erase(h->plans.begin());
    }
    h->plans.push_back(p);
    *out = p;
    return COVO_OK;
}

inline int hifi_enqueue(covo_hifigan* h, HifiPlan& p, const float* mel, void* wav, int out_dtype, cudaStream_t st) {
    const covo_hifigan_cfg& c = h->cfg;
    {
        ProfScope ps(PC_ELEMWISE, 0.0, st);
        dim3 g(ceil_div(p.T, 32), ceil_div(h->mel_pad, 32), p.B);
        mel_to_tc_kernel<<<g, 256, 0, st>>>(mel, p.mel_tc, c.num_mels, p.T, h->mel_pad, h->is_fp16);
        COVO_CK(cudaGetLastError());
    }
    COVO_TRY(launch_gemm(p.pre, st));
    for (int i = 0; i < c.num_upsamples; ++i) {
        HifiStage& s = p.stages[i];
        COVO_TRY(launch_gemm(s.up, st));
        if (p.fused_last && i + 1 == c.num_upsamples) {
            HifiFusedArgs f = p.fused;
            f.wav = wav;
            f.out_dtype = out_dtype;
            ProfScope ps(PC_GEMM_VOC, p.fused_flops, st);
            dim3 g(ceil_div(s.t_out, HF_TP), p.B);
            if (h->is_fp16) hifigan_fused_last_stage_kernel<true><<<g, HF_THREADS, HF_SMEM_BYTES, st>>>(f);
            else hifigan_fused_last_stage_kernel<false><<<g, HF_THREADS, HF_SMEM_BYTES, st>>>(f);
            COVO_CK(cudaGetLastError());
            return COVO_OK;
        }
        for (const GemmOp& op : s.convs) COVO_TRY(launch_gemm(op, st));
        const size_t n4 = static_cast<size_t>(p.B) * s.t_out * s.c_out_pad / 4;
        const float slope = (i + 1 == c.num_upsamples) ? 0.01f : 0.1f;     // models.py:103 vs :112 (F.leaky_relu default)
        ProfScope ps(PC_ELEMWISE, 0.0, st);
        stage_mean_act_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(
            s.xr[0], c.num_kernels > 1 ? s.xr[1] : nullptr, c.num_kernels > 2 ? s.xr[2] : nullptr, s.out_act, n4,
            1.0f / static_cast<float>(c.num_kernels), slope, h->is_fp16);
        COVO_CK(cudaGetLastError());
    }
    {
        const HifiStage& s = p.stages.back();
        Tensor tw, tb;
        COVO_TRY(h->w.get("conv_post.w", DT_F32, &tw));
        COVO_TRY(h->w.get("conv_post.b", DT_F32, &tb));
        ProfScope ps(PC_ELEMWISE, conv_flops(p.B, s.t_out, s.c_out, 1, 7), st);
        dim3 g(ceil_div(s.t_out, 256), p.B);
        conv_post_kernel<<<g, 256, 7 * s.c_out_pad * sizeof(float), st>>>(s.out_act, tw.as<float>(), tb.as<float>(), wav, s.t_out,
                                                                          s.c_out_pad, s.c_out_pad, h->is_fp16, out_dtype);
        COVO_CK(cudaGetLastError());
    }
    return COVO_OK;
}

}  // namespace covo
