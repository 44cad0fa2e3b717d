"""Checkpoint ingestion without pytorch_lightning / torch_ema / the reference package.

Acoustic model: a Lightning ``.ckpt`` is a ``torch.save`` dict with ``state_dict`` (keys ``cfm_wrapper.CoVoMix.*``),
``hyper_parameters`` (everything passed to ``CoVoMixModel.__init__`` -- including the ``data_module_cls`` CLASS OBJECT, so
plain unpickling needs ``covomix.data_module`` importable) and ``ema`` (``torch_ema`` state: ``shadow_params`` = list in
``self.parameters()`` order; covomix/conditional_model.py:192-201).  ``CoVoMixModel.eval()`` copies the EMA shadow
parameters over the live ones (conditional_model.py:203-217), so the inference weights are the shadow parameters, falling
back to ``state_dict`` when ``ema`` is missing (conditional_model.py:192-198).

Vocoder: ``{'generator': state_dict}`` next to ``vocoder_config.json`` (monologue_generation.py:368-386).
"""
from __future__ import annotations

import json
import os
import pickle
import types
import warnings
from typing import Dict, Tuple

import torch

from .packing import strip_flow_prefix
from .synthetic import HifiganConfig

# buffers of CoVoMix (everything else in its state_dict is a parameter, in named_parameters() order)
_FLOW_BUFFER_SUFFIXES = ("rotary_emb.inv_freq",)


class _Stub:
    """Placeholder for classes whose modules are not installed (Lightning, the reference's data module, ...)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state


def _stub_class(mod: str, name: str):
    return type(name, (_Stub,), {"__module__": mod})


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, mod, name):
        try:
            return super().find_class(mod, name)
        except (ImportError, AttributeError):
            return _stub_class(mod, name)


_tolerant_pickle = types.ModuleType("covomix_b200._tolerant_pickle")
_tolerant_pickle.Unpickler = _TolerantUnpickler
_tolerant_pickle.load = lambda f, **kw: _TolerantUnpickler(f, **kw).load()
_tolerant_pickle.__name__ = "pickle"


def _torch_load(path: str):
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_tolerant_pickle)


def load_acoustic_checkpoint(path: str, use_ema: bool = True) -> Tuple[Dict[str, torch.Tensor], dict]:
    """-> (CoVoMix state dict with bare keys, hyper_parameters as a plain dict of picklable values)."""
    ckpt = _torch_load(path)
    sd = strip_flow_prefix(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)
    ema = ckpt.get("ema") if isinstance(ckpt, dict) else None
    if use_ema and ema is not None and ema.get("shadow_params"):
        shadow = ema["shadow_params"]
        param_keys = [k for k in sd if not k.endswith(_FLOW_BUFFER_SUFFIXES)]
        if len(param_keys) != len(shadow):
            raise ValueError(f"EMA has {len(shadow)} shadow parameters but the state dict has {len(param_keys)} parameters")
        sd = dict(sd)
        for k, p in zip(param_keys, shadow):
            if tuple(p.shape) != tuple(sd[k].shape):
                raise ValueError(f"EMA shadow parameter for '{k}' has shape {tuple(p.shape)}, expected {tuple(sd[k].shape)}")
            sd[k] = p
    elif use_ema:
        warnings.warn("EMA state_dict not found in checkpoint; using the raw weights (as conditional_model.py:196-198)")
    hp = ckpt.get("hyper_parameters", {}) if isinstance(ckpt, dict) else {}
    hp = {k: v for k, v in dict(hp).items() if isinstance(v, (int, float, str, bool, type(None), list, tuple, dict))}
    return sd, hp


def load_vocoder_checkpoint(ckpt_path: str, config_path: str = None) -> Tuple[Dict[str, torch.Tensor], HifiganConfig]:
    """``ckpt_path``: the generator checkpoint; config defaults to ``vocoder_config.json`` in the same directory."""
    if config_path is None:
        config_path = os.path.join(os.path.dirname(os.path.abspath(ckpt_path)), "vocoder_config.json")
    with open(config_path) as f:
        cfg = HifiganConfig.from_json(json.load(f))
    ckpt = _torch_load(ckpt_path)
    sd = ckpt["generator"] if isinstance(ckpt, dict) and "generator" in ckpt else ckpt
    return sd, cfg
