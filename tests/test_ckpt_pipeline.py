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


# ---------------------------------------------------------------------------------------------- text-to-semantic host logic
def test_t2s_finish_targets_matches_oracle_loop_semantics():
    """What the reference's loop leaves in target / target2 (text2semantic.py:804-832), for every exit path."""
    from covomix_b200 import synthetic as syn
    from covomix_b200.t2s import finish_targets, set_eos_id
    from oracle import t2s_oracle as orc
    eos = syn.COMIX.semantic_eos_id
    tok = torch.tensor([[[5, eos, 7, 8], [9, 10, eos, 12]]])                 # [B=1, 2 streams, 4 steps]
    # two outputs, loop exhausted (no break): stream 1 masked after its EOS, stream 2 untouched
    t, m = finish_targets(tok, 4, False, syn.COMIX)
    assert t.tolist() == [[5, eos, -1, -1, 9, 10, eos, 12]] and m.tolist() == [[1, 1, 0, 0, 1, 1, 1, 1]]
    # two outputs, EOS rule fired: both masked
    t, _ = finish_targets(tok, 4, True, syn.COMIX)
    assert t.tolist() == [[5, eos, -1, -1, 9, 10, eos, -1]]
    # one output: masked only when the EOS rule fired
    t, _ = finish_targets(tok[:, :1], 4, False, syn.COSINGLE)
    assert t.tolist() == [[5, eos, 7, 8]]
    t, _ = finish_targets(tok[:, :1], 4, True, syn.COSINGLE)
    assert t.tolist() == [[5, eos, -1, -1]]
    ids = torch.tensor([[3, 4, 0, 0], [5, 6, 7, 8]])
    assert torch.equal(set_eos_id(ids.clone(), 99, 0), orc.set_eos_id(ids.clone(), 99, 0))
    assert set_eos_id(ids.clone(), 99, 0).tolist() == [[3, 4, 99, 0, 0], [5, 6, 7, 8, 99]]


def test_t2s_packing_layout_and_config_recovery():
    import numpy as np
    from covomix_b200 import packing, synthetic as syn
    for cfg in (syn.COSINGLE, syn.COMIX):
        sd = syn.synthetic_t2s_state_dict(cfg, 3)
        lightning = {"cfm_wrapper.model." + k: v for k, v in sd.items()}
        assert packing.t2s_config_from_state_dict(lightning) == cfg
        blob = packing.pack_t2s_weights(lightning, cfg)
        assert bytes(blob[:8]) == b"COVOWTS1"
        n = int(np.frombuffer(blob[8:12].tobytes(), dtype="<u4")[0])
        assert n == 5 + 9 * cfg.source_depth + 1 + 13 * cfg.target_depth
        fp32 = packing.pack_t2s_weights(sd, cfg, "fp32")
        assert fp32.nbytes > blob.nbytes


def test_dialogue_item_assembly():
    from covomix_b200.pipeline import dialogue_item
    it = dialogue_item(torch.arange(5), torch.arange(5) + 10, torch.ones(3, 160), torch.tensor([600, 2]), torch.tensor([7]))
    assert it["phoneme_ids"].tolist() == [[0, 10], [1, 11], [2, 12], [501, 7], [2, 157]]
    assert it["mask"].tolist() == [False, False, False, True, True]
    assert it["cond"].shape == (5, 160) and float(it["cond"][3:].abs().sum()) == 0.0


def test_accelerate_text2semantic_swaps_sample(monkeypatch):
    import types
    from covomix_b200 import dropin, synthetic as syn
    import covomix_b200.t2s as t2s_mod

    class FakeT2S:
        def __init__(self, sd, cfg, device, weight_format="bf16"):
            self.cfg = cfg

        def sample(self, grapheme_token_ids, temperature=1., cond_scale=1., beam_search_decode=False, prompt_mel=None):
            return torch.tensor([1, 2, 3])

    monkeypatch.setattr(t2s_mod, "B200TextToSemantic", FakeT2S)
    net = torch.nn.Module()
    sd = syn.synthetic_t2s_state_dict(syn.COMIX, 3)
    net.state_dict = lambda: sd
    net.parameters = lambda: iter([torch.zeros(1)])
    model = types.SimpleNamespace(cfm_wrapper=types.SimpleNamespace(model=net, sample=lambda **k: None))
    dropin.accelerate_text2semantic(model, device="cuda:0")
    assert model.cfm_wrapper._b200_t2s.cfg == syn.COMIX
    assert model.cfm_wrapper.sample(grapheme_token_ids=torch.zeros(1, 4)).tolist() == [1, 2, 3]


def test_covomix_dialogues_glue_with_fakes():
    """T2S -> id assembly -> acoustic -> vocoder glue (dialogue_generation.py:283-330) with stand-ins for the three models."""
    from covomix_b200 import pipeline

    class FakeT2S:
        def sample(self, ids, **kw):                      # two streams of len(text) tokens, flattened like the reference
            n = ids.shape[1]
            return torch.cat((torch.arange(n), torch.arange(n) + 100))

    class FakeSampler:
        device = torch.device("cpu")

        def sample(self, *, phoneme_ids, cond, mask, cond_scale, **kw):
            return phoneme_ids[..., 0:1].float().expand(-1, -1, 80).clone()

    class FakeGen:
        def __call__(self, mel, out_dtype="f32"):
            return (mel[:, 0, :] * 1.0).to(torch.int16)

    texts = [torch.arange(5), torch.arange(7), torch.arange(5)]
    prompts = [{"semantic_a": torch.zeros(4, dtype=torch.long), "semantic_b": torch.ones(4, dtype=torch.long),
                "mel": torch.zeros(3, 160)} for _ in texts]
    out = pipeline.covomix_dialogues(FakeT2S(), FakeSampler(), FakeGen(), texts, prompts, batch=2)
    assert [len(o) for o in out] == [5, 7, 5]             # generated frames only (mask), input order kept
    assert out[1].tolist() == list(range(7))
