"""Drop-in for the reference's entry scripts (monologue_generation.py / dialogue_generation.py).

The reference has no plugin layer on this path; the boundary is two Python call sites
(SURVEY.md section 8b):
    model.synthesis_sample(phoneme_ids=, cond=, mask=, cond_scale=)   conditional_model.py:295-302
        -> self.cfm_wrapper.sample(...)                                 acoustic.py:597-688
    generator(mel)                                                       hifi-gan/models.py:100-116
(and, one stage upstream, ``text2semantic.synthesis_sample_text2semantic(ids)`` -> ``cfm_wrapper.sample``,
conditional_model.py:313-321 -> text2semantic.py:1237-1251, swapped by ``accelerate_text2semantic``).
``accelerate_acoustic_model`` swaps ``model.cfm_wrapper.sample`` for the B200 sampler built from the
model's own (EMA-swapped, eval-mode) weights; ``accelerate_generator`` returns a callable with the
Generator's interface built from the live generator's state dict.  Scripts, checkpoints and tensor
shapes stay as they are.  See INTEGRATION.md for the two-line patch.
"""
from __future__ import annotations

import torch

from .flow import B200FlowSampler
from .packing import flow_config_from_state_dict, t2s_config_from_state_dict
from .synthetic import HifiganConfig
from .vocoder import B200Generator


def accelerate_acoustic_model(model, device=None, heads: int = 16, dim_head: int = 64):
    """``model``: a ``CoVoMixModel`` (or anything with ``.cfm_wrapper.CoVoMix`` and ``.cfm_wrapper.sample``)
    already in eval mode (so the EMA weights are the live ones, conditional_model.py:203-217)."""
    wrapper = model.cfm_wrapper
    net = wrapper.CoVoMix
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    if device is None:
        device = next(net.parameters()).device
    heads = getattr(net.transformer.layers[0][2], "heads", heads) if hasattr(net, "transformer") else heads
    cfg = flow_config_from_state_dict(sd, heads=heads, dim_head=dim_head)
    kw = getattr(wrapper, "odeint_kwargs", {"method": "midpoint", "options": {"step_size": 0.0625}})
    sampler = B200FlowSampler(sd, cfg, device, torchdiffeq_ode_method=kw["method"],
                              ode_step_size=kw["options"]["step_size"])
    wrapper._b200_sampler = sampler
    wrapper._reference_sample = wrapper.sample
    wrapper.sample = sampler.sample               # same keyword signature as acoustic.py:598-607
    return model


def accelerate_text2semantic(model, device=None, heads: int = 8, dim_head: int = 64, weight_format: str = "bf16"):
    """``model``: the text-to-semantic ``CoVoMixModel`` (``--t2s_ckpt``; ``.cfm_wrapper`` is a ``TextToSemanticWrapper``,
    conditional_model.py:136).  Swaps ``model.cfm_wrapper.sample`` -- what ``synthesis_sample_text2semantic``
    (conditional_model.py:313-321; ``comix_pred`` / ``cosingle_pred`` in the generation scripts) calls -- for
    ``B200TextToSemantic.sample`` (same keyword signature, same flattened id tensor)."""
    from .t2s import B200TextToSemantic
    wrapper = model.cfm_wrapper
    net = wrapper.model
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    if device is None:
        device = next(net.parameters()).device
    cfg = t2s_config_from_state_dict(sd, heads=heads, dim_head=dim_head)
    t2s = B200TextToSemantic(sd, cfg, device, weight_format=weight_format)
    wrapper._b200_t2s = t2s
    wrapper._reference_sample = wrapper.sample
    wrapper.sample = t2s.sample                   # text2semantic.py:1237-1251
    return model


def accelerate_generator(generator, h, device=None, h_format: str = "fp16"):
    """``generator``: the loaded ``Generator`` (weight norm removed or not); ``h``: its AttrDict / dict config."""
    hd = dict(h) if not isinstance(h, dict) else h
    cfg = HifiganConfig.from_json(hd)
    if device is None:
        device = next(generator.parameters()).device
    return B200Generator({k: v.detach() for k, v in generator.state_dict().items()}, cfg, device, h_format=h_format)


def mel_decode_to_wav(generator: B200Generator, mel: torch.Tensor):
    """monologue_generation.py:52-59 with the x32768 + int16 cast fused into the vocoder's last kernel."""
    return generator(mel, out_dtype="i16").reshape(-1).cpu().numpy()
