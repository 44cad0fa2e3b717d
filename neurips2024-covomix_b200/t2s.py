"""Host-side mirror of ``TextToSemantic.generate`` / ``TextToSemanticWrapper.sample``
(covomix/covomix_model/text2semantic.py:659-848, :1237-1251) backed by libcovomix_b200.so.

``B200TextToSemantic.sample(grapheme_token_ids, temperature=1., cond_scale=1., ...)`` has the signature the generation
scripts reach through ``CoVoMixModel.synthesis_sample_text2semantic`` (covomix/conditional_model.py:313-321;
``comix_pred`` dialogue_generation.py:332-344) and returns the same flattened, mask-selected id tensor.

The whole autoregressive loop is one persistent kernel; the host only (a) applies ``set_eos_id`` to the text ids, (b) draws
the Gumbel uniforms -- the reference draws them from the torch RNG inside its loop (``gumbel_noise`` :108-113), so they
are an input here, like ``y0`` of the flow sampler -- and (c) applies ``mask_after_eos`` to the result exactly as the
reference's loop does (:804-826).  Not supported (raise): classifier-free guidance (``cond_scale > 1``; the released
recipes train with ``cond_drop_prob = 0`` so the reference asserts too, :682), beam search, speculative decoding.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import _native as nat
from .packing import pack_t2s_weights, t2s_config_from_state_dict
from .synthetic import T2SConfig


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def set_eos_id(t: torch.Tensor, eos_id: int, pad_id: int) -> torch.Tensor:
    """text2semantic.py:57-66."""
    eos_idx = ((t == pad_id).cumsum(dim=-1) == 0).sum(dim=-1, keepdim=True).long()
    t = F.pad(t, (0, 1), value=pad_id)
    t[torch.arange(t.shape[0], device=t.device)[:, None], eos_idx] = eos_id
    return t


def mask_after_eos(target: torch.Tensor, eos_id: int, pad_id: int) -> torch.Tensor:
    """text2semantic.py:72-75."""
    m = (target == eos_id).cumsum(dim=-1) > 0
    m = F.pad(m, (1, -1), value=False)
    return target.masked_fill(m, pad_id)


def finish_targets(tokens: torch.Tensor, steps: int, stopped: bool, cfg: T2SConfig):
    """What the reference's loop leaves in ``target`` / ``target2`` (text2semantic.py:804-832): with two outputs stream 1
    is masked after its EOS on every iteration and stream 2 only when the EOS rule ended the loop; with one output the
    mask is applied only on that exit.  tokens [B, n_out, >=steps] -> (target [B, n_out*steps], target_mask)."""
    eos, pad = cfg.semantic_eos_id, cfg.semantic_pad_id
    streams = [tokens[:, s, :steps] for s in range(cfg.n_out)]
    if cfg.two_output:
        streams[0] = mask_after_eos(streams[0], eos, pad)
        if stopped:
            streams[1] = mask_after_eos(streams[1], eos, pad)
    elif stopped:
        streams[0] = mask_after_eos(streams[0], eos, pad)
    target = torch.cat(streams, dim=1)
    return target, target != pad


class B200TextToSemantic:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: Optional[T2SConfig] = None, device="cuda:0",
                 weight_format: str = "bf16", sm_limit: Optional[int] = None):
        self.cfg = cfg if cfg is not None else t2s_config_from_state_dict(state_dict)
        self.device = nat.resolve_device(device)
        if weight_format not in ("bf16", "fp32"):
            raise ValueError(weight_format)
        c = self.cfg
        ccfg = nat.T2SCfg(dim=c.dim, source_depth=c.source_depth, target_depth=c.target_depth, heads=c.heads,
                          dim_head=c.dim_head, num_text_token_ids=c.num_text_token_ids,
                          num_semantic_token_ids=c.num_semantic_token_ids, two_output=int(c.two_output),
                          target_transformer_dim=c.target_transformer_dim, ff_mult=c.ff_mult, text_pad_id=c.text_pad_id,
                          weight_format=nat.COVO_T2S_W_F32 if weight_format == "fp32" else nat.COVO_T2S_W_BF16)
        blob = pack_t2s_weights(state_dict, c, weight_format)
        self._h = C.c_void_p()
        nat.check(nat.lib().covo_t2s_create(C.byref(ccfg), blob.ctypes.data_as(C.c_void_p), blob.nbytes,
                                            self.device.index, C.byref(self._h)), "covo_t2s_create")
        if sm_limit:       # the decode kernel's grid: sm_limit CTAs instead of one per SM (stage overlap)
            nat.check(nat.lib().covo_t2s_set_sm_limit(self._h, int(sm_limit)), "covo_t2s_set_sm_limit")
        self._ws: Dict[tuple, torch.Tensor] = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            nat.lib().covo_t2s_destroy(self._h)
            self._h = C.c_void_p()
        self._ws = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, B: int, S: int, max_length: int) -> torch.Tensor:
        key = (B, S, max_length)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = nat.lib().covo_t2s_workspace_bytes(self._h, B, S, max_length)
            if nbytes == 0:
                raise ValueError(f"unsupported shape B={B}, S={S}, max_length={max_length}")
            if len(self._ws) >= 4:
                self._ws.pop(next(iter(self._ws)))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def draw_noise(self, B: int, max_length: int, reference_rng: bool = False) -> torch.Tensor:
        """Uniform(0,1) draws [max_length, n_out, B, n_logits].  ``reference_rng=True`` replays the reference's
        call pattern (one ``zeros_like(logits).uniform_(0, 1)`` per stream per step, text2semantic.py:108-110) so a
        seeded run consumes the device generator the same way; the default is a single draw (same distribution)."""
        c = self.cfg
        if not reference_rng:
            return torch.rand(max_length, c.n_out, B, c.n_logits, device=self.device)
        return torch.stack([torch.stack([torch.zeros(B, c.n_logits, device=self.device).uniform_(0, 1)
                                         for _ in range(c.n_out)]) for _ in range(max_length)])

    # --------------------------------------------------------------------------------
    @torch.inference_mode()
    def generate(self, source, *, source_type="text", target_type="speech", temperature=1., filter_thres=0.1,
                 source_mask=None, max_length=2048, beam_search_decode=False, spec_decode=False,
                 return_source=False, return_target_mask=False, cond_scale=1., prompt_mel=None,
                 noise: Optional[torch.Tensor] = None, forced: Optional[torch.Tensor] = None, return_debug=False,
                 ignore_eos=False):
        """``TextToSemantic.generate`` (text2semantic.py:659-848).  Extra keyword arguments (not in the reference):
        ``noise`` uniform draws [>=max_length, n_out, B, n_logits]; ``forced`` int64 [B, n_out, max_length] teacher
        forcing; ``ignore_eos`` never stops early (benchmarking); ``return_debug`` additionally returns a dict with per-step logits, the encoder output and the raw
        token / step counters."""
        if source_type != "text" or target_type != "speech":
            raise NotImplementedError("covomix_b200: only the text -> speech direction is on this path")
        if beam_search_decode or spec_decode:
            raise NotImplementedError("covomix_b200: beam search / speculative decoding are not on this path")
        if cond_scale != 1.:
            raise NotImplementedError("covomix_b200: classifier-free guidance for text-to-semantic is not supported "
                                      "(the reference asserts cond_drop_prob > 0 for it, text2semantic.py:682)")
        if source_mask is not None:
            raise NotImplementedError("covomix_b200: source_mask is derived from the pad id (text2semantic.py:726-727)")
        c = self.cfg
        src = source.to(device=self.device, dtype=torch.int64)
        if src.ndim != 2:
            raise ValueError(f"source must be [B, S] token ids, got {tuple(src.shape)}")
        B = src.shape[0]
        if B > 8:
            raise ValueError("covomix_b200: at most 8 rows per text-to-semantic call")
        ids = set_eos_id(src, c.text_eos_id, c.text_pad_id).contiguous()          # :716-721
        NB = 1 if B <= 1 else (2 if B <= 2 else (4 if B <= 4 else 8))
        S1 = ids.shape[1]
        if noise is None:
            noise = self.draw_noise(B, max_length)
        noise = noise.to(device=self.device, dtype=torch.float32)
        if noise.shape[0] < max_length or tuple(noise.shape[1:]) != (c.n_out, B, c.n_logits):
            raise ValueError(f"noise must be [>={max_length}, {c.n_out}, {B}, {c.n_logits}], got {tuple(noise.shape)}")
        noise = noise[:max_length]
        if forced is not None:
            forced = forced.to(device=self.device, dtype=torch.int64)
            if tuple(forced.shape) != (B, c.n_out, max_length):
                raise ValueError(f"forced must be {(B, c.n_out, max_length)}, got {tuple(forced.shape)}")
        if NB != B:            # the kernel takes 1, 2, 4 or 8 rows: replicate the last row (same noise -> same tokens)
            rep = NB - B
            ids = torch.cat((ids, ids[-1:].expand(rep, -1)), 0)
            noise = torch.cat((noise, noise[:, :, -1:].expand(-1, -1, rep, -1)), 2)
            if forced is not None:
                forced = torch.cat((forced, forced[-1:].expand(rep, -1, -1)), 0)
        ids, noise = ids.contiguous(), noise.contiguous()
        forced = None if forced is None else forced.contiguous()
        tokens = torch.zeros(NB, c.n_out, max_length, dtype=torch.int64, device=self.device)
        result = torch.zeros(4, dtype=torch.int32, device=self.device)
        logits = torch.zeros(max_length, c.n_out, NB, c.n_logits, device=self.device) if return_debug else None
        enc = torch.zeros(NB, S1, c.dim, device=self.device) if return_debug else None
        ws = self._workspace(NB, S1, max_length)
        k = math.ceil(filter_thres * c.n_logits)                                  # top_k :126-129
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib().covo_t2s_generate(self._h, _ptr(ids), _ptr(noise), _ptr(forced), _ptr(tokens), _ptr(result),
                                              _ptr(logits), _ptr(enc), NB, S1, max_length, float(temperature), k,
                                              nat.COVO_T2S_IGNORE_EOS if ignore_eos else 0,
                                              _ptr(ws), ws.numel(), C.c_void_p(stream)), "covo_t2s_generate")
        steps, stopped, aborted = (int(v) for v in result[:3].tolist())           # one D2H sync, as the reference's loop
        if aborted:
            raise RuntimeError("covomix_b200: the text-to-semantic decode kernel aborted (grid barrier timeout)")
        target, target_mask = finish_targets(tokens[:B], steps, bool(stopped), c)
        out = (target,)
        if return_source:
            out = (src,) + out
        if return_target_mask:
            out = out + (target_mask,)
        if return_debug:
            dbg = dict(steps=steps, stopped=bool(stopped), tokens=tokens[:B, :, :steps],
                       logits=logits[:steps, :, :B], enc=enc[:B])
            out = out + (dbg,)
        return out[0] if len(out) == 1 else out

    @torch.inference_mode()
    def sample(self, grapheme_token_ids, temperature=1., cond_scale=1., beam_search_decode=False, prompt_mel=None,
               **kw):
        """``TextToSemanticWrapper.sample`` (text2semantic.py:1237-1251): the non-masked part of the target."""
        target, target_mask = self.generate(grapheme_token_ids, source_type="text", target_type="speech",
                                            return_target_mask=True, return_source=False, temperature=temperature,
                                            beam_search_decode=beam_search_decode, cond_scale=cond_scale,
                                            prompt_mel=prompt_mel, **kw)
        return target[target_mask]

    def weight_bytes_per_step(self) -> int:
        return int(nat.lib().covo_t2s_weight_bytes_per_step(self._h))

    def launches_per_generate(self) -> int:
        return int(nat.lib().covo_t2s_launches_per_generate(self._h))
