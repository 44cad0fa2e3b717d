"""CPU tests of checkpoint ingestion (Lightning / torch_ema layout without those packages) and of the batching glue."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

import covomix_b200  # noqa: F401
from covomix_b200 import ckpt, pipeline, synthetic as syn

SMALL_FLOW = syn.FlowConfig(dim=128, depth=2, heads=2, dim_phoneme_emb=64)


def _write_lightning_ckpt(path, with_ema=True):
    sd = syn.synthetic_flow_state_dict(SMALL_FLOW, 3)
    # a hyper-parameter that is a CLASS OBJECT from a module that will not be importable at load time
    mod = types.ModuleType("covomix_fake_data_module")

    class SpecsDataModule:                                       # noqa: D401
        pass

    SpecsDataModule.__module__ = "covomix_fake_data_module"
    SpecsDataModule.__qualname__ = "SpecsDataModule"
    mod.SpecsDataModule = SpecsDataModule
    sys.modules["covomix_fake_data_module"] = mod
    state = {"cfm_wrapper.CoVoMix." + k: v for k, v in sd.items()}
    obj = {"state_dict": state, "hyper_parameters": {"lr": 1e-4, "CoVoMix_depth": 2, "data_module_cls": SpecsDataModule},
           "epoch": 3}
    shadow = None
    if with_ema:
        shadow = [v + 1.0 for k, v in sd.items() if not k.endswith("inv_freq")]
        obj["ema"] = {"decay": 0.999, "num_updates": 10, "shadow_params": shadow, "collected_params": None}
    torch.save(obj, path)
    del sys.modules["covomix_fake_data_module"]
    return sd, shadow


def test_lightning_checkpoint_without_lightning(tmp_path):
    p = str(tmp_path / "acoustic.ckpt")
    sd, shadow = _write_lightning_ckpt(p, with_ema=True)
    with pytest.raises(Exception):
        torch.load(p, map_location="cpu", weights_only=False)     # plain unpickling needs the missing module
    got, hp = ckpt.load_acoustic_checkpoint(p)
    assert set(got) == set(sd) and hp == {"lr": 1e-4, "CoVoMix_depth": 2}
    # inference weights = EMA shadow parameters (conditional_model.py:203-217); buffers untouched
    assert torch.equal(got["to_embed.weight"], sd["to_embed.weight"] + 1.0)
    assert torch.equal(got["transformer.rotary_emb.inv_freq"], sd["transformer.rotary_emb.inv_freq"])
    raw, _ = ckpt.load_acoustic_checkpoint(p, use_ema=False)
    assert torch.equal(raw["to_embed.weight"], sd["to_embed.weight"])


def test_checkpoint_without_ema_falls_back(tmp_path):
    p = str(tmp_path / "acoustic_noema.ckpt")
    sd, _ = _write_lightning_ckpt(p, with_ema=False)
    with pytest.warns(UserWarning):
        got, _ = ckpt.load_acoustic_checkpoint(p)
    assert torch.equal(got["to_pred.weight"], sd["to_pred.weight"])


def test_vocoder_checkpoint_with_weight_norm(tmp_path):
    from covomix_b200 import packing
    cfg = syn.HifiganConfig(upsample_rates=(4, 2), upsample_kernel_sizes=(8, 4), upsample_initial_channel=16,
                            resblock_kernel_sizes=(3,), resblock_dilation_sizes=((1, 3),), num_mels=8)
    sd = syn.synthetic_hifigan_state_dict(cfg, 2)
    wn = {}
    for k, v in sd.items():                                       # re-express every weight as weight_g / weight_v
        if k.endswith(".weight"):
            g = v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.ndim - 1)))
            wn[k[:-7] + ".weight_g"], wn[k[:-7] + ".weight_v"] = g, v * 3.0
        else:
            wn[k] = v
    torch.save({"generator": wn}, str(tmp_path / "g_00400000"))
    with open(tmp_path / "vocoder_config.json", "w") as f:
        json.dump({"resblock": "1", "upsample_rates": [4, 2], "upsample_kernel_sizes": [8, 4], "upsample_initial_channel": 16,
                   "resblock_kernel_sizes": [3], "resblock_dilation_sizes": [[1, 3]], "num_mels": 8}, f)
    got, gcfg = ckpt.load_vocoder_checkpoint(str(tmp_path / "g_00400000"))
    assert gcfg == cfg
    folded = packing.fold_weight_norm(got)
    for k, v in sd.items():
        assert torch.allclose(folded[k], v, atol=1e-6), k


class _FakeSampler:
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def sample(self, *, phoneme_ids, cond, mask=None, cond_scale=1.0, y0=None):
        self.calls.append(tuple(cond.shape))
        assert phoneme_ids.shape[:2] == cond.shape[:2] == mask.shape
        return cond[:, :, :80] + phoneme_ids.reshape(cond.shape[0], cond.shape[1], -1)[:, :, :1].float()


class _FakeGenerator:
    def __init__(self):
        self.calls = []

    def __call__(self, mel, out_dtype="f32"):
        self.calls.append(tuple(mel.shape))
        assert out_dtype == "i16" and mel.ndim == 3 and mel.shape[1] == 80
        return (mel.sum(1, keepdim=True).repeat_interleave(4, dim=2)).to(torch.int16)


def test_synthesize_groups_equal_lengths_and_keeps_order():
    g = torch.Generator().manual_seed(0)
    items, lens, prompts = [], [40, 24, 40, 40, 24, 33], [8, 8, 8, 10, 8, 5]
    for n, p in zip(lens, prompts):
        mask = torch.zeros(n, dtype=torch.bool)
        mask[p:] = True
        items.append({"phoneme_ids": torch.randint(0, 500, (n,), generator=g), "cond": torch.randn(n, 80, generator=g), "mask": mask})
    smp, gen = _FakeSampler(), _FakeGenerator()
    wavs = pipeline.synthesize(smp, gen, items, batch=2, vocoder_batch=4)
    assert sorted(smp.calls) == sorted([(2, 40, 80), (1, 40, 80), (2, 24, 80), (1, 33, 80)])
    assert sorted(c[2] for c in gen.calls) == sorted([32, 30, 16, 28])       # T = N - prompt; equal T batched together
    for it, w, n, p in zip(items, wavs, lens, prompts):
        mel = (it["cond"] + it["phoneme_ids"][:, None].float())[p:].t()
        expect = mel.sum(0, keepdim=True).repeat_interleave(4, dim=1).to(torch.int16).numpy().reshape(-1)
        assert w.dtype == np.int16 and np.array_equal(w, expect)
    assert pipeline.concat_turns([wavs[1], wavs[4]]).shape[0] == wavs[1].shape[0] + wavs[4].shape[0]
