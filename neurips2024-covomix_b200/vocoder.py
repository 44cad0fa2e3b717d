"""Host-side mirror of the HiFi-GAN ``Generator`` (hifi-gan/models.py:75-125 ==
covomix/vocoder/models.py) backed by libcovomix_b200.so.  ``gen(mel)`` takes the reference's input
([80, T] or [B, 80, T] fp32) and returns the reference's output shape ([1, L] / [B, 1, L] fp32,
L = 160*T + 32 for config_covomix.json); ``eval()`` / ``remove_weight_norm()`` / ``to()`` exist so the
generation scripts' load sequence (monologue_generation.py:382-386) runs unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _native as nat
from .packing import pack_hifigan_weights
from .synthetic import HIFIGAN_COVOMIX, HifiganConfig


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class B200Generator:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: HifiganConfig = HIFIGAN_COVOMIX, device="cuda:0",
                 h_format: str = "fp16", sm_limit=None):
        self.h = cfg
        self.device = nat.resolve_device(device)
        nk = len(cfg.resblock_kernel_sizes)
        nd = len(cfg.resblock_dilation_sizes[0])
        ccfg = nat.HifiganCfg()
        ccfg.num_mels = cfg.num_mels
        ccfg.upsample_initial_channel = cfg.upsample_initial_channel
        ccfg.num_upsamples = len(cfg.upsample_rates)
        for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
            ccfg.upsample_rates[i] = u
            ccfg.upsample_kernel_sizes[i] = k
        ccfg.num_kernels = nk
        ccfg.num_dilations = nd
        for j in range(nk):
            ccfg.resblock_kernel_sizes[j] = cfg.resblock_kernel_sizes[j]
            for m in range(nd):
                ccfg.resblock_dilations[j][m] = cfg.resblock_dilation_sizes[j][m]
        ccfg.resblock_type = int(cfg.resblock)
        ccfg.h_format = nat.COVO_H_FP16 if h_format == "fp16" else nat.COVO_H_BF16
        blob = pack_hifigan_weights(state_dict, cfg, h_format)
        self._h = C.c_void_p()
        nat.check(nat.lib().covo_hifigan_create(C.byref(ccfg), blob.ctypes.data_as(C.c_void_p), blob.nbytes,
                                                self.device.index, C.byref(self._h)), "covo_hifigan_create")
        if sm_limit:
            nat.check(nat.lib().covo_hifigan_set_sm_limit(self._h, int(sm_limit)), "covo_hifigan_set_sm_limit")
        self._ws: Dict[tuple, torch.Tensor] = {}

    # reference-API no-ops (weights are already folded / on device)
    def eval(self):
        return self

    def remove_weight_norm(self):
        return None

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("covomix_b200 has no CPU path")
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            nat.lib().covo_hifigan_destroy(self._h)
            self._h = C.c_void_p()
        self._ws = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def out_len(self, T: int) -> int:
        return int(nat.lib().covo_hifigan_out_len(self._h, T))

    def launches_per_forward(self) -> int:
        return nat.lib().covo_hifigan_launches_per_forward(self._h)

    def _workspace(self, B: int, T: int) -> torch.Tensor:
        key = (B, T)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = nat.lib().covo_hifigan_workspace_bytes(self._h, B, T)
            while len(self._ws) >= 2:                       # keep at most two call shapes resident
                self._ws.pop(next(iter(self._ws)))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # receptive field of the generator is ~20.4 mel frames per side; chunks of a long input are
    # computed independently with this halo and stitched (exact away from the sequence ends, see the locality test)
    HALO = 24
    MAX_FRAMES_PER_CALL = 16384                             # B*T per library call: bounds the workspace to ~6 GB

    def _forward_one(self, x: torch.Tensor, code: int, tdt) -> torch.Tensor:
        B, _, T = x.shape
        wav = torch.empty(B, 1, self.out_len(T), dtype=tdt, device=self.device)
        ws = self._workspace(B, T)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib().covo_hifigan_forward(self._h, _ptr(x), _ptr(wav), B, T, code, _ptr(ws), ws.numel(),
                                                 C.c_void_p(stream)), "covo_hifigan_forward")
        return wav

    @torch.inference_mode()
    def forward(self, mel: torch.Tensor, out_dtype: str = "f32") -> torch.Tensor:
        unbatched = mel.ndim == 2
        x = mel[None] if unbatched else mel
        if x.ndim != 3 or x.shape[1] != self.h.num_mels:
            raise ValueError(f"mel must be [{self.h.num_mels}, T] or [B, {self.h.num_mels}, T], got {tuple(mel.shape)}")
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        B, _, T = x.shape
        code, tdt = {"f32": (nat.COVO_WAV_F32, torch.float32), "f16": (nat.COVO_WAV_F16, torch.float16),
                     "i16": (nat.COVO_WAV_I16, torch.int16)}[out_dtype]
        if unbatched:
            # Generator.forward on [80, T]: Conv1d treats it as an unbatched [C, T] input, so the result is [1, L]
            # (hifi-gan/models.py:100-116) -- not [1, 1, L]
            return self.forward(x, out_dtype)[0]
        if B * T <= self.MAX_FRAMES_PER_CALL:
            return self._forward_one(x, code, tdt)
        hop, halo = self.h.hop, self.HALO
        tc = max(4 * halo, self.MAX_FRAMES_PER_CALL // B)            # frames produced per chunk
        if tc >= T:                                                  # a single item is too long only through B: split batch
            return torch.cat([self.forward(x[b:b + 1], out_dtype) for b in range(B)], dim=0)
        wav = torch.empty(B, 1, self.out_len(T), dtype=tdt, device=self.device)
        for s in range(0, T, tc):
            e = min(T, s + tc)
            s0, e0 = max(0, s - halo), min(T, e + halo)
            part = self._forward_one(x[:, :, s0:e0].contiguous(), code, tdt)
            lo = (s - s0) * hop
            hi = part.shape[-1] if e == T else (e - s0) * hop
            wav[:, :, s * hop:s * hop + (hi - lo)] = part[:, :, lo:hi]
        return wav

    __call__ = forward
