// Host-side plumbing shared by the flow sampler and the vocoder: error reporting, packed-weight blob,
// TMA tensor-map encoding, GEMM op construction/launch.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/covomix_b200.h"
#include "attention_sm100.cuh"
#include "gemm_sm100.cuh"
#include "kernels.cuh"

namespace covo {

// ------------------------------------------------------------------------------------------------ errors
inline std::string& err_slot() {
    static thread_local std::string s;
    return s;
}
inline int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err_slot() = buf;
    return code;
}
#define COVO_CK(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return ::covo::fail(COVO_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define COVO_TRY(expr)              \
    do {                            \
        int rc_ = (expr);           \
        if (rc_ != COVO_OK) return rc_; \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ------------------------------------------------------------------------------------------------ weights blob
// Layout (little endian), written by covomix_b200.packing:
//   header : char magic[8] = "COVOWTS1"; u32 n_entries; u32 reserved
//   entries: n x { char name[48]; u32 dtype; u32 ndim; u64 shape[4]; u64 offset; u64 nbytes }
//   data   : tensors at `offset` bytes from the blob start (256-byte aligned)
enum : uint32_t { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2, DT_I64 = 3 };
struct BlobEntry {
    char name[48];
    uint32_t dtype, ndim;
    uint64_t shape[4];
    uint64_t offset, nbytes;
};
struct Tensor {
    void* ptr = nullptr;
    uint32_t dtype = 0, ndim = 0;
    uint64_t shape[4] = {0, 0, 0, 0};
    template <class T>
    T* as() const { return static_cast<T*>(ptr); }
};
struct Weights {
    void* dev = nullptr;
    size_t bytes = 0;
    std::map<std::string, Tensor> t;

    int load(const void* host_blob, size_t nbytes) {
        if (nbytes < 16 || memcmp(host_blob, "COVOWTS1", 8) != 0) return fail(COVO_ERR_WEIGHTS, "bad weight blob magic");
        const uint8_t* p = static_cast<const uint8_t*>(host_blob);
        uint32_t n;
        memcpy(&n, p + 8, 4);
        if (16 + static_cast<size_t>(n) * sizeof(BlobEntry) > nbytes) return fail(COVO_ERR_WEIGHTS, "weight blob truncated");
        COVO_CK(cudaMalloc(&dev, nbytes));
        COVO_CK(cudaMemcpy(dev, host_blob, nbytes, cudaMemcpyHostToDevice));
        bytes = nbytes;
        for (uint32_t i = 0; i < n; ++i) {
            BlobEntry e;
            memcpy(&e, p + 16 + static_cast<size_t>(i) * sizeof(BlobEntry), sizeof(BlobEntry));
            if (e.offset + e.nbytes > nbytes || (e.offset & 255)) return fail(COVO_ERR_WEIGHTS, "entry %u out of range", i);
            Tensor tt;
            tt.ptr = static_cast<uint8_t*>(dev) + e.offset;
            tt.dtype = e.dtype;
            tt.ndim = e.ndim;
            memcpy(tt.shape, e.shape, sizeof(tt.shape));
            e.name[47] = 0;
            t[e.name] = tt;
        }
        return COVO_OK;
    }
    int get(const std::string& name, uint32_t dtype, Tensor* out) const {
        auto it = t.find(name);
        if (it == t.end()) return fail(COVO_ERR_WEIGHTS, "packed weights: tensor '%s' missing", name.c_str());
        if (it->second.dtype != dtype)
            return fail(COVO_ERR_WEIGHTS, "packed weights: tensor '%s' has dtype %u, expected %u", name.c_str(),
                        it->second.dtype, dtype);
        *out = it->second;
        return COVO_OK;
    }
    void release() {
        if (dev) cudaFree(dev);
        dev = nullptr;
    }
};

// ------------------------------------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// Innermost dim contiguous; strides (bytes) for dims 1..rank-1; 128B swizzle.  dt: 0 = bf16, 1 = fp16, 2 = fp32.
inline int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int dt) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(COVO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) {
            gstr[i - 1] = strides_bytes[i - 1];
            if (gstr[i - 1] % 16) return fail(COVO_ERR_INVALID, "TMA stride %llu not 16-byte aligned", (unsigned long long)gstr[i - 1]);
        }
    }
    if (reinterpret_cast<uintptr_t>(base) % 16) return fail(COVO_ERR_INVALID, "TMA base not 16-byte aligned");
    const CUtensorMapDataType cdt = dt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                    : (dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    CUresult r = fn(tm, cdt, rank,
                    const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(COVO_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return COVO_OK;
}

// ------------------------------------------------------------------------------------------------ launch profiler
// Opt-in (covo_prof_begin/_end): brackets every kernel launch with CUDA events on the launching stream so that
// bench.py can report per-kernel-class time shares and the dominant kernel's achieved FLOP/s.  Off in normal use.
enum : int { PC_GEMM = 0, PC_ATTN = 1, PC_NORM = 2, PC_CONVPOS = 3, PC_ELEMWISE = 4, PC_PROLOGUE = 5, PC_GEMM_VOC = 6, PC_T2S_DECODE = 7, PC_FLOW_PERSISTENT = 8, PC_COUNT = 9 };
struct ProfRec {
    int cat;
    double flops;
    cudaEvent_t e0, e1;
};
struct Profiler {
    bool on = false;
    std::vector<ProfRec> recs;
};
inline Profiler& prof() {
    static Profiler p;
    return p;
}
struct ProfScope {
    cudaStream_t st;
    int idx = -1;
    ProfScope(int cat, double flops, cudaStream_t s) : st(s) {
        Profiler& p = prof();
        if (!p.on) return;
        ProfRec r;
        r.cat = cat;
        r.flops = flops;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        p.recs.push_back(r);
        idx = static_cast<int>(p.recs.size()) - 1;
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(prof().recs[idx].e1, st);
    }
};

// ------------------------------------------------------------------------------------------------ GEMM ops
struct DeviceInfo {
    int device = 0;
    int num_sms = 148;
    int gemm_mc = 1;         // 2: GEMMs with >= 2 M tiles run as cluster pairs sharing the weight tile by TMA multicast
    int gemm_cg = 0;         // 2: ... as cta_group::2 pairs (one M = 256 MMA per pair, each CTA stages half of the weight tile);
                             // 0 (default): pairs for GEMMs of at least two waves of tiles, independent CTAs for small ones; 1: never
};

struct ASource {             // 16-bit activations [Z][rows][K], K contiguous
    const void* ptr;
    int K;                   // multiple of 64 (padded columns are zero)
    int rows;
    int Z;
    long long row_stride;    // elements
    long long z_stride;      // elements
};

struct GemmOp {
    GemmArgs args;
    int bn = 256;
    int fmt = 1;             // 0 fp16, 1 bf16
    int grid = 1;
    int mc = 1;              // 2: cluster-pair kernel (gemm_tc_pair_kernel), tmB box is (64, bn / 2)
    int cg = 1;              // 2: cta_group::2 kernel (gemm_tc_cg2_kernel), tmB box is (64, bn / 2); excludes mc = 2
    int cat = PC_GEMM;       // profiler class
    double flops = 0.0;      // algorithmic FLOPs of this launch (real channels only; padding does not count)
};

template <int BN, int FMT>
inline int set_gemm_attr() {
    static bool done = false;
    if (!done) {
        COVO_CK(cudaFuncSetAttribute(gemm_tc_kernel<BN, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     GemmCfg<BN>::SMEM_BYTES));
        COVO_CK(cudaFuncSetAttribute(gemm_tc_pair_kernel<BN, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     GemmCfg<BN>::SMEM_BYTES));
        if (BN >= 128)
            COVO_CK(cudaFuncSetAttribute(gemm_tc_cg2_kernel<(BN >= 128 ? BN : 128), FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GemmCfg<(BN >= 128 ? BN : 128), 2>::SMEM_BYTES));
        done = true;
    }
    return COVO_OK;
}

// Tile width: minimise  waves * cost(BN)  with waves = ceil(tiles / SMs).  cost grows sub-linearly with 1/BN because
// narrow tiles re-read the A operand from smem more often per FLOP (BN = 64 runs the tensor pipe at ~2/3 rate; measured
// in profiles/r01_gemm_microbench.txt), so e.g. M = 1300, N = 1024 takes one wave of 88 BN=128 tiles rather than two
// waves of 176 BN=64 tiles.
inline int pick_bn(int n_pad, long long m_tiles, int num_sms, int force_bn) {
    if (force_bn && n_pad % force_bn == 0) return force_bn;
    const int cands[3] = {256, 128, 64};
    const double cost[3] = {256.0, 128.0 * 1.10, 64.0 * 1.50};
    int best = 0;
    double best_t = 0.0;
    for (int i = 0; i < 3; ++i) {
        if (n_pad % cands[i]) continue;
        const long long tiles = m_tiles * (n_pad / cands[i]);
        const double t = static_cast<double>((tiles + num_sms - 1) / num_sms) * cost[i];
        if (best == 0 || t < best_t) {
            best = cands[i];
            best_t = t;
        }
    }
    return best;
}

// Fills tensor maps, tile counts and grid.  Epilogue/output-mapping fields of op.args must be set by the caller
// (before or after; this function touches only tmA/tmB/rows/Z/n_tiles/taps/kc_per_tap and bn/grid/fmt).
inline int build_gemm(GemmOp& op, const DeviceInfo& di, const ASource& a, int q_rows, int q_Z, const void* w, int n_pad,
                      int taps, int is_fp16, int force_bn = 0) {
    if (a.K % GEMM_BK) return fail(COVO_ERR_INVALID, "GEMM K=%d not a multiple of %d", a.K, GEMM_BK);
    if (taps < 1 || taps > GEMM_MAX_TAPS) return fail(COVO_ERR_INVALID, "GEMM taps=%d out of range", taps);
    const long long m_tiles = static_cast<long long>(q_Z) * ceil_div(q_rows, GEMM_BM);
    op.bn = pick_bn(n_pad, m_tiles, di.num_sms, force_bn);
    if (op.bn == 0 || n_pad % op.bn) return fail(COVO_ERR_INVALID, "GEMM N_pad=%d not tileable", n_pad);
    op.fmt = is_fp16 ? 0 : 1;
    const int m_tiles_z = ceil_div(q_rows, GEMM_BM);
    const bool cg_wanted = di.gemm_cg == 2 || (di.gemm_cg == 0 && di.gemm_mc != 2 && m_tiles * (n_pad / op.bn) >= 2LL * di.num_sms);
    op.cg = (cg_wanted && m_tiles_z >= 2 && di.num_sms >= 2 && op.bn >= 128) ? 2 : 1;
    op.mc = (op.cg == 1 && di.gemm_mc == 2 && m_tiles_z >= 2 && di.num_sms >= 2) ? 2 : 1;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.rows), static_cast<uint64_t>(a.Z)};
        uint64_t str[2] = {static_cast<uint64_t>(a.row_stride) * 2, static_cast<uint64_t>(a.z_stride) * 2};
        if (a.Z == 1) str[1] = static_cast<uint64_t>(a.row_stride) * 2 * static_cast<uint64_t>(a.rows > 0 ? a.rows : 1);
        uint32_t box[3] = {GEMM_BK, GEMM_BM, 1};
        COVO_TRY(make_tmap(&op.args.tmA, a.ptr, 3, dims, str, box, is_fp16));
    }
    {
        const int ktot = taps * a.K;
        uint64_t dims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(n_pad)};
        uint64_t str[1] = {static_cast<uint64_t>(ktot) * 2};
        uint32_t box[2] = {GEMM_BK, static_cast<uint32_t>(op.bn / (op.mc * op.cg))};     // pair kernels: each CTA fetches half of the tile
        COVO_TRY(make_tmap(&op.args.tmB, w, 2, dims, str, box, is_fp16));
    }
    op.args.rows = q_rows;
    op.args.Z = q_Z;
    op.args.n_tiles = n_pad / op.bn;
    op.args.taps = taps;
    op.args.kc_per_tap = a.K / GEMM_BK;
    if (op.mc == 2 || op.cg == 2) {
        const long long pairs = static_cast<long long>(q_Z) * ceil_div(m_tiles_z, 2) * op.args.n_tiles;
        const long long clusters = di.num_sms / 2;
        op.grid = 2 * static_cast<int>(pairs < clusters ? pairs : clusters);
        return COVO_OK;
    }
    const long long tiles = m_tiles * op.args.n_tiles;
    op.grid = static_cast<int>(tiles < di.num_sms ? tiles : di.num_sms);
    if (op.grid < 1) op.grid = 1;
    return COVO_OK;
}

inline void gemm_defaults(GemmArgs& g) {
    memset(&g, 0, sizeof(g));
    g.up_s = 1;
    g.up_p = 0;
    g.phase_w = 1 << 30;
    g.slope = 0.1f;
}

// Box-shaped outputs [Z][rows][ld] (n contiguous): TMA maps for the fp32 output, the fp32 residual (may be the same
// buffer) and the 16-bit output.  n_valid columns and `rows` rows are stored; everything else is clipped by TMA.
inline int gemm_set_outputs(GemmOp& op, float* out_f32, const float* residual, void* out_h, int n_valid, int rows, int Z,
                            long long ld, long long zs, int h_is_fp16) {
    GemmArgs& g = op.args;
    if (residual != nullptr && out_f32 == nullptr) return fail(COVO_ERR_INVALID, "GEMM residual requires an fp32 output");
    g.n_valid = n_valid;
    g.scatter = 0;
    g.has_out_f32 = out_f32 != nullptr;
    g.has_residual = residual != nullptr;
    g.has_out_h = out_h != nullptr;
    g.h_is_fp16 = h_is_fp16;
    g.out_f32 = out_f32;
    g.out_h = out_h;
    if (Z == 1) zs = ld * rows;
    uint64_t dims[3] = {static_cast<uint64_t>(n_valid), static_cast<uint64_t>(rows), static_cast<uint64_t>(Z)};
    if (out_f32 != nullptr || residual != nullptr) {
        uint64_t str[2] = {static_cast<uint64_t>(ld) * 4, static_cast<uint64_t>(zs) * 4};
        uint32_t box[3] = {32, 32, 1};
        if (out_f32 != nullptr) COVO_TRY(make_tmap(&g.tmOutF, out_f32, 3, dims, str, box, 2));
        if (residual != nullptr) COVO_TRY(make_tmap(&g.tmRes, residual, 3, dims, str, box, 2));
    }
    if (out_h != nullptr) {
        uint64_t str[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(zs) * 2};
        uint32_t box[3] = {64, 32, 1};
        COVO_TRY(make_tmap(&g.tmOutH, out_h, 3, dims, str, box, h_is_fp16 ? 1 : 0));
    }
    return COVO_OK;
}

inline int launch_gemm(const GemmOp& op, cudaStream_t st) {
    // the 16-bit output is packed in the operands' format (a compile-time constant of the kernel)
    if (op.args.has_out_h && !op.args.scatter && (op.args.h_is_fp16 != 0) != (op.fmt == 0))
        return fail(COVO_ERR_INVALID, "GEMM 16-bit output format differs from the operand format");
    ProfScope ps(op.cat, op.flops, st);
#define COVO_LAUNCH(BN_, F_)                                                                              \
    do {                                                                                                  \
        COVO_TRY((set_gemm_attr<BN_, F_>()));                                                             \
        if (op.cg == 2) gemm_tc_cg2_kernel<(BN_ >= 128 ? BN_ : 128), F_><<<op.grid, GEMM_THREADS, GemmCfg<(BN_ >= 128 ? BN_ : 128), 2>::SMEM_BYTES, st>>>(op.args); \
        else if (op.mc == 2) gemm_tc_pair_kernel<BN_, F_><<<op.grid, GEMM_THREADS, GemmCfg<BN_>::SMEM_BYTES, st>>>(op.args); \
        else gemm_tc_kernel<BN_, F_><<<op.grid, GEMM_THREADS, GemmCfg<BN_>::SMEM_BYTES, st>>>(op.args);   \
    } while (0)
    if (op.fmt == 1) {
        if (op.bn == 256) COVO_LAUNCH(256, 1);
        else if (op.bn == 128) COVO_LAUNCH(128, 1);
        else COVO_LAUNCH(64, 1);
    } else {
        if (op.bn == 256) COVO_LAUNCH(256, 0);
        else if (op.bn == 128) COVO_LAUNCH(128, 0);
        else COVO_LAUNCH(64, 0);
    }
#undef COVO_LAUNCH
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

// ------------------------------------------------------------------------------------------------ attention launch
// Tunables of attention_tc_kernel (defaults chosen from tools/micro/attn_bench.cu on B200, profiles/r02_attention_bench.txt;
// overridable for A/B runs):
//   COVO_ATT_POLY    0: every exponential on the MUFU, 1 (default): one pair in four on the FMA pipe, 2: three pairs in eight
//   COVO_ATT_TOKEN   1: exponential token between the two query groups (anti-phase); default 0 (free running) -- the token
//                    costs as much as it hides while one warp per sub-partition cannot saturate the MUFU on its own
struct AttnTune {
    int poly = 1;
    int stagger = 0;
};
inline AttnTune& attn_tune() {
    static AttnTune t;
    static bool init = false;
    if (!init) {
        if (const char* v = getenv("COVO_ATT_POLY")) t.poly = atoi(v);
        if (const char* v = getenv("COVO_ATT_TOKEN")) t.stagger = atoi(v);
        init = true;
    }
    return t;
}
// HANDOFF = 0 compiles the exponential-token code out (the default: args.stagger == 0).  Left in as run-time tests on a uniform
// flag it cost 9 % of the kernel (289 -> 263 us at the C3 shape): four extra branches per key tile with one softmax warp per
// sub-partition and group, nothing to hide them behind.
template <int MASK>
inline int attn_set_attrs_mask() {
    COVO_CK(cudaFuncSetAttribute(attention_tc_kernel<MASK, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    COVO_CK(cudaFuncSetAttribute(attention_tc_kernel<MASK, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    return COVO_OK;
}
inline int attn_set_attrs() {
    COVO_TRY(attn_set_attrs_mask<0>());
    COVO_TRY(attn_set_attrs_mask<0x88>());
    COVO_TRY(attn_set_attrs_mask<0x92>());
    return COVO_OK;
}
// qkv: bf16 [Bt, N, 3*heads*64] (the to_qkv output, RoPE applied); out: bf16 [Bt, N, heads*64]
inline int attn_build_args(AttnArgs& a, const void* qkv, void* out, int Bt, int N, int heads) {
    const int inner = heads * ATT_D;
    uint64_t dims[3] = {static_cast<uint64_t>(3 * inner), static_cast<uint64_t>(N), static_cast<uint64_t>(Bt)};
    uint64_t str[2] = {static_cast<uint64_t>(3 * inner) * 2, static_cast<uint64_t>(3 * inner) * 2 * N};
    uint32_t box[3] = {64, 128, 1};
    COVO_TRY(make_tmap(&a.tmQKV, qkv, 3, dims, str, box, 0));
    a.out = static_cast<__nv_bfloat16*>(out);
    a.N = N;
    a.heads = heads;
    a.inner = inner;
    a.stagger = attn_tune().stagger;
    a.reverse = 0;
    attn_fill_items(a, Bt);
    return COVO_OK;
}
template <int MASK>
inline void launch_attention_mask(const AttnArgs& a, int grid, cudaStream_t st) {
    if (a.stagger != 0) attention_tc_kernel<MASK, 128><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
    else attention_tc_kernel<MASK, 0><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
}
inline int launch_attention_kernel(const AttnArgs& a, int num_sms, cudaStream_t st) {
    const int grid = a.n_items < num_sms ? a.n_items : num_sms;
    switch (attn_tune().poly) {
        case 0: launch_attention_mask<0>(a, grid, st); break;
        case 2: launch_attention_mask<0x92>(a, grid, st); break;
        default: launch_attention_mask<0x88>(a, grid, st); break;
    }
    COVO_CK(cudaGetLastError());
    return COVO_OK;
}

// bump allocator over the caller's workspace
struct Arena {
    uint8_t* base;
    size_t cap, off = 0;
    Arena(void* p, size_t c) : base(static_cast<uint8_t*>(p)), cap(c) {}
    template <class T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base ? base + off : nullptr);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

inline int check_device(int device, DeviceInfo* di) {
    cudaDeviceProp p;
    COVO_CK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return fail(COVO_ERR_ARCH, "device %d is sm_%d%d; this library is sm_100a only", device, p.major, p.minor);
    di->device = device;
    di->num_sms = p.multiProcessorCount;
    return COVO_OK;
}

}  // namespace covo
