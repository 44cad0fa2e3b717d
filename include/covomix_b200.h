/* covomix_b200.h -- C ABI of libcovomix_b200.so (CUDA, sm_100a only).
 *
 * The reference (vivian556123/NeurIPS2024-CoVoMix) is pure Python on this path and has no FFI layer;
 * the drop-in boundary is two Python call sites.  Each entry point below cites the reference
 * interface it replaces (paths relative to the reference root).  The Python binding a maintainer
 * would add is in INTEGRATION.md; the binding shipped here is neurips2024-covomix_b200/_native.py.
 *
 * Conventions: every function returns 0 on success and a negative covo_status on failure; the
 * message is available from covo_last_error() (thread-local).  Nothing throws across the ABI.
 * Pointers are DEVICE pointers unless marked host.  Calls only enqueue work on `stream`
 * (a cudaStream_t passed as void*); they never synchronise the device (the one-time build of a plan
 * for a new (B, N, solver) shape does host work and a stream capture).  A handle is bound to one
 * device and is not re-entrant; use one handle per rank.  The caller owns inputs, outputs and the
 * workspace; the library owns only its copy of the packed weights.
 */
#ifndef COVOMIX_B200_H
#define COVOMIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COVO_ABI_VERSION 1

typedef enum covo_status {
    COVO_OK = 0,
    COVO_ERR_INVALID = -1,   /* bad argument / shape / config */
    COVO_ERR_CUDA = -2,      /* a CUDA runtime or driver call failed */
    COVO_ERR_WEIGHTS = -3,   /* packed weight blob malformed or a tensor is missing */
    COVO_ERR_ARCH = -4       /* device is not sm_100 (no fallback path exists) */
} covo_status;

enum { COVO_ODE_EULER = 0, COVO_ODE_MIDPOINT = 1 };            /* torchdiffeq method= (acoustic.py:572,589) */
enum { COVO_WAV_F32 = 0, COVO_WAV_F16 = 1, COVO_WAV_I16 = 2 }; /* waveform output dtype */
enum { COVO_H_BF16 = 0, COVO_H_FP16 = 1 };                     /* 16-bit operand format of the vocoder convs */

/* ---- flow-matching acoustic decoder -------------------------------------------------------------
 * Mirrors the constructor arguments of CoVoMix (covomix/covomix_model/acoustic.py:326-348) as set by
 * CoVoMixModel.__init__ (covomix/conditional_model.py:99-115). */
typedef struct covo_flow_cfg {
    int32_t dim;                 /* 1024 */
    int32_t depth;               /* 8 (even) */
    int32_t heads;               /* 16 */
    int32_t dim_head;            /* 64 (only value supported) */
    int32_t dim_in;              /* 80 VoSingle, 160 VoMix: width of `cond` */
    int32_t dim_x;               /* 80: width of the ODE state / to_pred output */
    int32_t n_streams;           /* 1: ids [B,N]; 2: ids [B,N,2] (twocondition_oneoutput) */
    int32_t num_phoneme_tokens;  /* 502; id 502 is the CFG null token (acoustic.py:367) */
    int32_t dim_phoneme_emb;     /* 1024 */
    int32_t ff_mult;             /* 4 */
    int32_t conv_pos_kernel;     /* 31 */
} covo_flow_cfg;

typedef struct covo_flow covo_flow;

/* packed_weights: HOST pointer to a blob produced by covomix_b200.packing.pack_flow_weights. */
int covo_flow_create(const covo_flow_cfg* cfg, const void* packed_weights, size_t bytes, int device, covo_flow** out);
int covo_flow_destroy(covo_flow* h);

/* Bytes of caller-owned scratch needed by covo_flow_sample / covo_flow_velocity for batch B, length N. */
size_t covo_flow_workspace_bytes(const covo_flow* h, int B, int N, int n_eval_times);

/* Replaces ConditionalFlowMatcherWrapper.sample (acoustic.py:597-688) == CoVoMixModel.synthesis_sample
 * (conditional_model.py:295-302) with torchdiffeq.odeint's fixed-grid solver inlined:
 *   ids   int64 [B,N] or [B,N,2]        cond f32 [B,N,dim_in]
 *   y0    f32 [B,N,dim_x]  (the caller draws it with torch.randn_like, acoustic.py:647-650)
 *   out   f32 [B,N,dim_x]  x(t=1), prompt rows included (the caller slices with its mask)
 *   method/n_steps: COVO_ODE_MIDPOINT,16 is the reference default (step_size 0.0625, acoustic.py:568)
 *   cond_scale: classifier-free guidance scale (acoustic.py:421-428); 1.0 skips the null branch. */
int covo_flow_sample(covo_flow* h, const int64_t* ids, const float* cond, const float* y0, float* out, int B, int N,
                     int method, int n_steps, float cond_scale, void* workspace, size_t workspace_bytes, void* stream);

/* torchdiffeq's fixed grid (FixedGridODESolver, options={'step_size': h}, acoustic.py:586-591) is k*h with the last point
 * snapped to 1, so a step size that does not divide 1 (e.g. 0.3 -> 0, .3, .6, .9, 1) ends on a shorter step.  After this
 * call covo_flow_sample builds that grid; its n_steps argument must then equal ceil(1/step_size).  0 restores k/n_steps. */
int covo_flow_set_step_size(covo_flow* h, float step_size);

/* One velocity evaluation  v = CoVoMix.forward_with_cond_scale(x, times=t, ...) (acoustic.py:414-428):
 * x, v f32 [B,N,dim_x].  Used by parity tests and for the "ODE-step ms" metric. */
int covo_flow_velocity(covo_flow* h, const int64_t* ids, const float* cond, const float* x, float t, float* v, int B,
                       int N, float cond_scale, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels one covo_flow_sample call launches for this configuration (bench.py's gpu_launches). */
int covo_flow_launches_per_sample(const covo_flow* h, int method, int n_steps, float cond_scale);

/* Kernels launched by the most recent covo_flow_sample call on this handle.  With COVO_FLOW_PERSISTENT=1 (opt-in: measured
 * slower than the CUDA graph) every evaluation and solver update of the call runs inside ONE persistent
 * cooperative kernel (csrc/flow_persistent.cuh);
 * the count is then 4 (embedding gather, e_const GEMM, state -> input, that kernel) instead of covo_flow_launches_per_sample. */
int covo_flow_last_launches(const covo_flow* h);

/* ---- HiFi-GAN generator ---------------------------------------------------------------------------
 * Mirrors the fields Generator.__init__ reads from the JSON config (hifi-gan/models.py:76-98,
 * hifi-gan/config_covomix.json). */
typedef struct covo_hifigan_cfg {
    int32_t num_mels;             /* 80 */
    int32_t upsample_initial_channel;
    int32_t num_upsamples;        /* <= 8 */
    int32_t upsample_rates[8];
    int32_t upsample_kernel_sizes[8];
    int32_t num_kernels;          /* resblocks per stage, <= 4 */
    int32_t resblock_kernel_sizes[4];
    int32_t num_dilations;        /* convs per resblock, <= 4 */
    int32_t resblock_dilations[4][4];
    int32_t resblock_type;        /* 1 = ResBlock1 (models.py:11-48), 2 = ResBlock2 (models.py:51-72) */
    int32_t h_format;             /* COVO_H_BF16 | COVO_H_FP16 */
} covo_hifigan_cfg;

typedef struct covo_hifigan covo_hifigan;

/* packed_weights: HOST pointer to a blob produced by covomix_b200.packing.pack_hifigan_weights
 * (weight norm already folded: Generator.remove_weight_norm, models.py:118-125). */
int covo_hifigan_create(const covo_hifigan_cfg* cfg, const void* packed_weights, size_t bytes, int device,
                        covo_hifigan** out);
int covo_hifigan_destroy(covo_hifigan* h);
size_t covo_hifigan_workspace_bytes(const covo_hifigan* h, int B, int T);
/* Output samples per item for T mel frames (160*T + 32 for config_covomix.json). */
int64_t covo_hifigan_out_len(const covo_hifigan* h, int T);

/* Replaces Generator.forward (hifi-gan/models.py:100-116 == covomix/vocoder/models.py):
 *   mel f32 [B, num_mels, T]  ->  wav [B, out_len(T)] in out_dtype (COVO_WAV_*);
 *   COVO_WAV_I16 additionally applies mel_decode_to_wav's x32768 + int16 cast
 *   (monologue_generation.py:52-59). */
int covo_hifigan_forward(covo_hifigan* h, const float* mel, void* wav, int B, int T, int out_dtype, void* workspace,
                         size_t workspace_bytes, void* stream);
int covo_hifigan_launches_per_forward(const covo_hifigan* h);

/* ---- text-to-semantic (CoSingle / CoMix) -----------------------------------------------------------------
 * Mirrors the arguments of TextToSemantic(...) (covomix/covomix_model/text2semantic.py:417-451) as passed by
 * CoVoMixModel.__init__ (covomix/conditional_model.py:122-135; running_command/T2S_Co{Single,Mix}.sh). */
enum { COVO_T2S_W_BF16 = 0, COVO_T2S_W_F32 = 1 };   /* storage type of the decoder's matrices (accumulation is fp32) */
typedef struct covo_t2s_cfg {
    int32_t dim;                     /* 512: source transformer / cross-attention context width */
    int32_t source_depth;            /* 4 */
    int32_t target_depth;            /* 4 (<= 8) */
    int32_t heads;                   /* 8 */
    int32_t dim_head;                /* 64 (only value supported) */
    int32_t num_text_token_ids;      /* 30530; id 30530 is the text EOS (text2semantic.py:491-494) */
    int32_t num_semantic_token_ids;  /* 501;   id 501 is the semantic EOS */
    int32_t two_output;              /* 0 CoSingle, 1 CoMix (two token streams per step, :768-776) */
    int32_t target_transformer_dim;  /* 512 CoSingle, 1024 CoMix */
    int32_t ff_mult;                 /* 4 */
    int32_t text_pad_id;             /* 0 */
    int32_t weight_format;           /* COVO_T2S_W_* */
} covo_t2s_cfg;

typedef struct covo_t2s covo_t2s;

/* packed_weights: HOST pointer to a blob produced by covomix_b200.packing.pack_t2s_weights. */
int covo_t2s_create(const covo_t2s_cfg* cfg, const void* packed_weights, size_t bytes, int device, covo_t2s** out);
int covo_t2s_destroy(covo_t2s* h);
size_t covo_t2s_workspace_bytes(const covo_t2s* h, int B, int S, int max_length);

/* Replaces TextToSemantic.generate(source, source_type='text', target_type='speech', ...) (text2semantic.py:659-848; the
 * non-beam, non-speculative branch; cond_scale == 1) == TextToSemanticWrapper.sample (:1237-1251):
 *   text_ids   int64 [B,S]   text token ids AFTER set_eos_id (:57-66, done by the host mirror), pad = text_pad_id
 *   u          f32   [max_length, n_out, B, n_logits]  uniform(0,1) draws of gumbel_noise (:108-113) in the reference's
 *              draw order (the caller draws them with torch, as for y0 of the flow sampler); n_logits = 502
 *   forced     int64 [B, n_out, max_length] or NULL: teacher forcing -- token fed back at each position (parity tests)
 *   tokens     int64 [B, n_out, max_length]  sampled ids (positions >= steps are not written)
 *   result     int32 [4] (device): [0] decoding steps executed, [1] 1 if the EOS rule ended the loop (:804-826),
 *              [2] 1 if the kernel aborted on a stuck grid barrier
 *   logits_out f32 [max_length, n_out, B, n_logits] or NULL: pre-filter logits of every step
 *   enc_out    f32 [B, S, dim] or NULL: output of the source transformer (parity tests)
 *   top_k: ceil(0.1 * n_logits) = 51 for filter_logits_fn = top_k (:126-132).
 *   flags: COVO_T2S_IGNORE_EOS = never stop early (benchmarking on random weights, where EOS would end the loop at a
 *          random position); 0 = the reference's rule. */
enum { COVO_T2S_IGNORE_EOS = 1 };
int covo_t2s_generate(covo_t2s* h, const int64_t* text_ids, const float* u, const int64_t* forced, int64_t* tokens,
                      int32_t* result, float* logits_out, float* enc_out, int B, int S, int max_length, float temperature,
                      int top_k, int flags, void* workspace, size_t workspace_bytes, void* stream);
/* Kernels launched by one covo_t2s_generate call (the whole autoregressive loop is one of them). */
int covo_t2s_launches_per_generate(const covo_t2s* h);
/* Bytes of decoder matrices one decoding step streams (the quantity the step is bound by). */
size_t covo_t2s_weight_bytes_per_step(const covo_t2s* h);

/* ---- prompt mel front-end -----------------------------------------------------------------------------------
 * Replaces mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)
 * (covomix/util/generate_mel.py:49-72), which extract_mel / prepare_oracle_hubert call on the acoustic prompt
 * (monologue_generation.py:62-90) with the constants of monologue_generation.py:349-357. */
typedef struct covo_mel_cfg {
    int32_t n_fft;      /* 480 */
    int32_t hop_size;   /* 160 */
    int32_t win_size;   /* 480 (<= n_fft) */
    int32_t num_mels;   /* 80 */
} covo_mel_cfg;
typedef struct covo_mel covo_mel;
/* window: HOST f32 [win_size] (torch.hann_window); mel_basis: HOST f32 [num_mels, n_fft/2 + 1]
 * (librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax); covomix_b200.frontend.slaney_mel_filterbank restates it). */
int covo_mel_create(const covo_mel_cfg* cfg, const float* window, const float* mel_basis, int device, covo_mel** out);
int covo_mel_destroy(covo_mel* h);
/* Frames for L samples: (L + 2*pad - n_fft) / hop + 1 with pad = (n_fft - hop) / 2 (L >= pad + 1). */
int covo_mel_frames(const covo_mel* h, int L);
/* wav f32 [B, L] in [-1, 1] -> mel f32 [B, num_mels, frames(L)] = log(clamp(mel_basis @ |STFT|, 1e-5)). */
int covo_mel_forward(covo_mel* h, const float* wav, float* mel, int B, int L, void* stream);

/* ---- SM budgets (stage overlap) ----------------------------------------------------------------------------------
 * Every hot kernel is persistent (grid = min(work, SMs)).  Limiting a handle to n_sms lets two stages share the GPU on
 * two streams without time-slicing each other -- e.g. the text-to-semantic loop of the next batch on 28 SMs
 * next to the flow sampler of the current batch on the other 120 (bench.py --workload c4p).  Call before the first
 * sample / forward / generate of the handle (plans built earlier keep their grids).  n_sms <= 0 restores the device count. */
int covo_flow_set_sm_limit(covo_flow* h, int n_sms);
int covo_hifigan_set_sm_limit(covo_hifigan* h, int n_sms);
int covo_t2s_set_sm_limit(covo_t2s* h, int n_sms);

/* ---- misc ------------------------------------------------------------------------------------------ */
const char* covo_last_error(void);
int covo_version(void);

/* ---- launch profiler (bench.py's roofline leg) -----------------------------------------------------------
 * Between covo_prof_begin() and covo_prof_end() every kernel launch is bracketed by CUDA events on its stream
 * (CUDA graphs are bypassed).  Classes: 0 tcgen05 GEMM (flow), 1 attention, 2 RMSNorm/AdaLN, 3 conv-pos, 4 element-wise,
 * 5 per-call prologue, 6 tcgen05 GEMM launched by the vocoder.  covo_prof_end synchronises the device and returns, per class, the summed kernel time (ms),
 * the summed algorithmic FLOPs and the launch count. */
int covo_prof_begin(void);
int covo_prof_end(double* ms_per_class, double* flops_per_class, int* launches_per_class, int n_classes);

/* ---- kernel-level test hooks (used only by tests/ to check single kernels against torch) -------------
 * D[M,N] = A[M,K] * W[N,K]^T (+bias) (+residual), bf16 operands, fp32 accumulate, via the tcgen05 kernel. */
int covo_dbg_gemm(const void* A_bf16, const void* W_bf16, const float* bias, const float* residual, float* out_f32,
                  void* out_bf16, int M, int N, int K, int act_h, int force_bn, void* stream);
/* qkv bf16 [Bt, N, 3*heads*64] -> out bf16 [Bt, N, heads*64]; impl 0 = tcgen05 kernel, 1 = naive CUDA-core kernel. */
int covo_dbg_attention(const void* qkv_bf16, void* out_bf16, int Bt, int N, int heads, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COVOMIX_B200_H */
