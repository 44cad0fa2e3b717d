"""Host-side mirror of ``ConditionalFlowMatcherWrapper`` (covomix/covomix_model/acoustic.py:560-688)
backed by libcovomix_b200.so.  Same call signature and return shape as the reference's
``sample`` (and therefore ``CoVoMixModel.synthesis_sample``, covomix/conditional_model.py:295-302);
the solver settings that the reference hard-codes in the constructor (acoustic.py:566-574:
``torchdiffeq_ode_method='midpoint'``, ``ode_step_size=0.0625``) are constructor arguments here too.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _native as nat
from .packing import flow_config_from_state_dict, pack_flow_weights
from .synthetic import FlowConfig


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class B200FlowSampler:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: Optional[FlowConfig] = None, device="cuda:0",
                 torchdiffeq_ode_method: str = "midpoint", ode_step_size: float = 0.0625, sm_limit: Optional[int] = None,
                 validate_ids: bool = True):
        self.cfg = cfg if cfg is not None else flow_config_from_state_dict(state_dict)
        self.device = nat.resolve_device(device)
        self.method = torchdiffeq_ode_method
        self.step_size = float(ode_step_size)
        if not 0.0 < self.step_size <= 1.0:
            raise ValueError(f"ode_step_size must be in (0, 1], got {ode_step_size}")
        self.validate_ids = bool(validate_ids)     # one min/max reduction + host read per call (the reference's IndexError)
        c = self.cfg
        ccfg = nat.FlowCfg(dim=c.dim, depth=c.depth, heads=c.heads, dim_head=c.dim_head, dim_in=c.dim_in, dim_x=c.dim_x,
                           n_streams=c.n_streams, num_phoneme_tokens=c.num_phoneme_tokens,
                           dim_phoneme_emb=c.dim_phoneme_emb, ff_mult=c.ff_mult, conv_pos_kernel=c.conv_pos_kernel)
        blob = pack_flow_weights(state_dict, c)
        self._h = C.c_void_p()
        L = nat.lib()
        nat.check(L.covo_flow_create(C.byref(ccfg), blob.ctypes.data_as(C.c_void_p), blob.nbytes,
                                     self.device.index, C.byref(self._h)), "covo_flow_create")
        # torchdiffeq's grid is k*h with the last point snapped to 1 (a step size that does not divide 1 ends on a shorter
        # step); the library builds exactly that grid from the step size
        nat.check(L.covo_flow_set_step_size(self._h, self.step_size), "covo_flow_set_step_size")
        if sm_limit:       # persistent kernels of this handle use at most sm_limit SMs (stage overlap, see include/covomix_b200.h)
            nat.check(L.covo_flow_set_sm_limit(self._h, int(sm_limit)), "covo_flow_set_sm_limit")
        self._ws: Dict[tuple, torch.Tensor] = {}

    # --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            nat.lib().covo_flow_destroy(self._h)
            self._h = C.c_void_p()
        self._ws = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, B: int, N: int, n_t: int) -> torch.Tensor:
        key = (B, N, n_t)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = nat.lib().covo_flow_workspace_bytes(self._h, B, N, n_t)
            if len(self._ws) >= 4:
                self._ws.pop(next(iter(self._ws)))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def _check_inputs(self, phoneme_ids, cond):
        c = self.cfg
        if cond.ndim != 3 or cond.shape[-1] != c.dim_in:
            raise ValueError(f"cond must be [B, N, {c.dim_in}], got {tuple(cond.shape)}")
        B, N, _ = cond.shape
        want = (B, N, 2) if c.n_streams == 2 else (B, N)
        if tuple(phoneme_ids.shape) != want:
            raise ValueError(f"phoneme_ids must be {want}, got {tuple(phoneme_ids.shape)}")
        ids = phoneme_ids.to(device=self.device, dtype=torch.int64).contiguous()
        cond = cond.to(device=self.device, dtype=torch.float32).contiguous()
        if self.validate_ids and ids.numel():
            # nn.Embedding(num_phoneme_tokens + 1, ...) raises on ids outside the table (acoustic.py:367-368, :496-500)
            lo, hi = (int(v) for v in torch.stack((ids.min(), ids.max())).tolist())
            if lo < 0 or hi > c.num_phoneme_tokens:
                raise IndexError(f"phoneme_ids out of range: [{lo}, {hi}] not within [0, {c.num_phoneme_tokens}]")
        return ids, cond, B, N

    def n_steps(self) -> int:
        """Number of solver steps of torchdiffeq's fixed grid: ceil(1/h) evaluated in fp32 like the library does."""
        one, h = torch.tensor(1.0), torch.tensor(self.step_size, dtype=torch.float32)
        return int(math.ceil(float(one / h)))

    # --------------------------------------------------------------------------------
    @torch.inference_mode()
    def sample(self, *, phoneme_ids, cond, mask=None, steps=3, cond_scale=1., decode_to_audio=False, y0=None):
        """``ConditionalFlowMatcherWrapper.sample`` (acoustic.py:597-688).  ``mask`` is accepted and ignored
        exactly like the reference (it never reaches the network, acoustic.py:627-633); ``steps`` only
        selects output times there and the last one (t=1) is returned, so it has no effect either.
        ``y0`` (not in the reference signature) lets tests supply the noise; default is
        ``torch.randn_like`` on the device as in acoustic.py:647-650."""
        ids, cond, B, N = self._check_inputs(phoneme_ids, cond)
        c = self.cfg
        if y0 is None:
            y0 = torch.randn_like(cond if c.n_streams == 1 else cond[:, :, :80])
        y0 = y0.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(y0.shape) != (B, N, c.dim_x):
            raise ValueError(f"y0 must be {(B, N, c.dim_x)}, got {tuple(y0.shape)}")
        method = {"euler": nat.COVO_ODE_EULER, "midpoint": nat.COVO_ODE_MIDPOINT}[self.method]
        n_steps = self.n_steps()
        n_t = n_steps * (2 if method == nat.COVO_ODE_MIDPOINT else 1)
        ws = self._workspace(B, N, n_t)
        out = torch.empty(B, N, c.dim_x, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib().covo_flow_sample(self._h, _ptr(ids), _ptr(cond), _ptr(y0), _ptr(out), B, N, method, n_steps,
                                             float(cond_scale), _ptr(ws), ws.numel(), C.c_void_p(stream)),
                  "covo_flow_sample")
        return out

    @torch.inference_mode()
    def velocity(self, x, *, times, phoneme_ids, cond, cond_scale=1.):
        """``CoVoMix.forward_with_cond_scale`` (acoustic.py:414-428) for a scalar time."""
        ids, cond, B, N = self._check_inputs(phoneme_ids, cond)
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        ws = self._workspace(B, N, 1)
        v = torch.empty(B, N, self.cfg.dim_x, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib().covo_flow_velocity(self._h, _ptr(ids), _ptr(cond), _ptr(x), float(times), _ptr(v), B, N,
                                               float(cond_scale), _ptr(ws), ws.numel(), C.c_void_p(stream)),
                  "covo_flow_velocity")
        return v

    def last_launches(self) -> int:
        """Kernels launched by the most recent ``sample`` call (4 on the persistent path, else launches_per_sample)."""
        return nat.lib().covo_flow_last_launches(self._h)

    def launches_per_sample(self, cond_scale: float = 0.7) -> int:
        method = {"euler": nat.COVO_ODE_EULER, "midpoint": nat.COVO_ODE_MIDPOINT}[self.method]
        return nat.lib().covo_flow_launches_per_sample(self._h, method, self.n_steps(), float(cond_scale))
