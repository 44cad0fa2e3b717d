"""ctypes binding of libcovomix_b200.so (C ABI declared in include/covomix_b200.h).

There is no CPU or PyTorch fallback: if the library has not been built, or the device is not
sm_100, every entry point raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
(one nvcc invocation on ``csrc/covomix_b200.cu``, flags in ``__graft_entry__.NVCC_FLAGS``).
"""
from __future__ import annotations

import ctypes as C
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libcovomix_b200.so")

COVO_ODE_EULER, COVO_ODE_MIDPOINT = 0, 1
COVO_WAV_F32, COVO_WAV_F16, COVO_WAV_I16 = 0, 1, 2
COVO_H_BF16, COVO_H_FP16 = 0, 1

EXPORTS = (
    "covo_flow_create", "covo_flow_destroy", "covo_flow_workspace_bytes", "covo_flow_sample", "covo_flow_velocity",
    "covo_flow_launches_per_sample", "covo_hifigan_create", "covo_hifigan_destroy", "covo_hifigan_workspace_bytes",
    "covo_hifigan_out_len", "covo_hifigan_forward", "covo_hifigan_launches_per_forward", "covo_last_error",
    "covo_version", "covo_dbg_gemm", "covo_dbg_attention", "covo_prof_begin", "covo_prof_end",
    "covo_t2s_create", "covo_t2s_destroy", "covo_t2s_workspace_bytes", "covo_t2s_generate",
    "covo_t2s_launches_per_generate", "covo_t2s_weight_bytes_per_step",
    "covo_mel_create", "covo_mel_destroy", "covo_mel_frames", "covo_mel_forward",
    "covo_flow_set_sm_limit", "covo_hifigan_set_sm_limit", "covo_t2s_set_sm_limit", "covo_flow_set_step_size", "covo_flow_last_launches",
)


class FlowCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "dim", "depth", "heads", "dim_head", "dim_in", "dim_x", "n_streams", "num_phoneme_tokens", "dim_phoneme_emb",
        "ff_mult", "conv_pos_kernel")]


class HifiganCfg(C.Structure):
    _fields_ = [
        ("num_mels", C.c_int32), ("upsample_initial_channel", C.c_int32), ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8), ("num_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * 4), ("num_dilations", C.c_int32),
        ("resblock_dilations", (C.c_int32 * 4) * 4), ("resblock_type", C.c_int32), ("h_format", C.c_int32),
    ]


class T2SCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "dim", "source_depth", "target_depth", "heads", "dim_head", "num_text_token_ids", "num_semantic_token_ids",
        "two_output", "target_transformer_dim", "ff_mult", "text_pad_id", "weight_format")]


class MelCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_fft", "hop_size", "win_size", "num_mels")]


COVO_T2S_W_BF16, COVO_T2S_W_F32 = 0, 1
COVO_T2S_IGNORE_EOS = 1

_lib = None


def resolve_device(device):
    """``torch.device`` with an explicit CUDA ordinal: ``"cuda"`` without an index means the CURRENT device (what
    ``torch.cuda.set_device(local_rank)`` selected), not GPU 0.  There is no CPU path."""
    import torch
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("covomix_b200 has no CPU path; device must be a CUDA (sm_100) device")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"covomix_b200: native library {LIB_PATH} is missing -- build it first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, sz, i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64
    L.covo_last_error.restype = C.c_char_p
    L.covo_version.restype = i32
    L.covo_flow_create.argtypes = [C.POINTER(FlowCfg), vp, sz, i32, C.POINTER(vp)]
    L.covo_flow_destroy.argtypes = [vp]
    L.covo_flow_workspace_bytes.argtypes = [vp, i32, i32, i32]
    L.covo_flow_workspace_bytes.restype = sz
    L.covo_flow_sample.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, sz, vp]
    L.covo_flow_velocity.argtypes = [vp, vp, vp, vp, f32, vp, i32, i32, f32, vp, sz, vp]
    L.covo_flow_launches_per_sample.argtypes = [vp, i32, i32, f32]
    L.covo_flow_set_step_size.argtypes = [vp, f32]
    L.covo_flow_last_launches.argtypes = [vp]
    L.covo_hifigan_create.argtypes = [C.POINTER(HifiganCfg), vp, sz, i32, C.POINTER(vp)]
    L.covo_hifigan_destroy.argtypes = [vp]
    L.covo_hifigan_workspace_bytes.argtypes = [vp, i32, i32]
    L.covo_hifigan_workspace_bytes.restype = sz
    L.covo_hifigan_out_len.argtypes = [vp, i32]
    L.covo_hifigan_out_len.restype = i64
    L.covo_hifigan_forward.argtypes = [vp, vp, vp, i32, i32, i32, vp, sz, vp]
    L.covo_hifigan_launches_per_forward.argtypes = [vp]
    L.covo_dbg_gemm.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.covo_dbg_attention.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.covo_t2s_create.argtypes = [C.POINTER(T2SCfg), vp, sz, i32, C.POINTER(vp)]
    L.covo_t2s_destroy.argtypes = [vp]
    L.covo_t2s_workspace_bytes.argtypes = [vp, i32, i32, i32]
    L.covo_t2s_workspace_bytes.restype = sz
    L.covo_t2s_generate.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, i32, vp, sz, vp]
    L.covo_t2s_launches_per_generate.argtypes = [vp]
    L.covo_t2s_weight_bytes_per_step.argtypes = [vp]
    L.covo_t2s_weight_bytes_per_step.restype = sz
    L.covo_mel_create.argtypes = [C.POINTER(MelCfg), vp, vp, i32, C.POINTER(vp)]
    L.covo_mel_destroy.argtypes = [vp]
    L.covo_mel_frames.argtypes = [vp, i32]
    L.covo_mel_forward.argtypes = [vp, vp, vp, i32, i32, vp]
    for fn in (L.covo_flow_set_sm_limit, L.covo_hifigan_set_sm_limit, L.covo_t2s_set_sm_limit):
        fn.argtypes = [vp, i32]
    L.covo_prof_begin.argtypes = []
    L.covo_prof_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), i32]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("covo_version",):
            fn.restype = i32
    _lib = L
    return L


PROF_CLASSES = ("gemm_tc", "attention_tc", "rmsnorm", "convpos", "elementwise", "prologue", "gemm_tc_vocoder",
                "t2s_decode", "flow_persistent")


class profile:
    """Context manager around covo_prof_begin/_end; ``.result`` maps class -> (ms, flops, launches)."""

    def __enter__(self):
        check(lib().covo_prof_begin(), "covo_prof_begin")
        self.result = {}
        return self

    def __exit__(self, *exc):
        n = len(PROF_CLASSES)
        ms, fl, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_int * n)()
        check(lib().covo_prof_end(ms, fl, cnt, n), "covo_prof_end")
        self.result = {PROF_CLASSES[i]: (ms[i], fl[i], cnt[i]) for i in range(n)}
        return False


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().covo_last_error()
        raise RuntimeError(f"covomix_b200: {what} failed (status {rc}): {msg.decode() if msg else ''}")
