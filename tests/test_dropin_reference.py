"""Drop-in plumbing exercised on the REAL reference objects (CPU; needs /root/reference, so it runs in the build
container and is skipped on the GPU box, where only the golden vectors travel).

The reference modules are imported read-only with the same stubs ``tests/golden/make_golden.py`` uses (matplotlib,
torchode, torchdiffeq -- none of them carries arithmetic on this path).  What is checked: configuration recovery from
the live module's ``state_dict`` (``flow_config_from_state_dict`` / ``t2s_config_from_state_dict`` /
``HifiganConfig.from_json(AttrDict)``), recovery of the solver settings from ``odeint_kwargs`` (acoustic.py:586-591),
packing of the live weights into the library's blob, and that the swapped call sites keep the reference's keyword
signatures (acoustic.py:598-607, text2semantic.py:1237, models.py:100).  The native sampler / generator classes are
replaced by recorders: there is no GPU here and no CPU fallback to fall back on.
"""
import inspect
import json
import os
import sys

import numpy as np
import pytest
import torch

import covomix_b200  # noqa: F401
from covomix_b200 import dropin, packing, synthetic as syn

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "covomix")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_modules():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg
    mg._install_stubs()
    sys.path.insert(0, REF)
    from covomix.covomix_model import acoustic, text2semantic
    sys.path.insert(0, os.path.join(REF, "hifi-gan"))
    import env
    import models
    return acoustic, text2semantic, models, env


class Recorder:
    """Stands in for B200FlowSampler / B200TextToSemantic / B200Generator: records constructor arguments."""
    made = []

    def __init__(self, sd, cfg, device, **kw):
        self.sd, self.cfg, self.device, self.kw = sd, cfg, device, kw
        Recorder.made.append(self)

    def sample(self, *a, **k):
        return ("b200", a, k)

    def __call__(self, mel):
        return ("b200", mel)


class Shim:
    """The two attributes of CoVoMixModel that the path touches (conditional_model.py:99-136, :295-321)."""

    def __init__(self, wrapper):
        self.cfm_wrapper = wrapper

    def synthesis_sample(self, phoneme_ids, cond, mask, cond_scale):
        return self.cfm_wrapper.sample(phoneme_ids=phoneme_ids, cond=cond, mask=mask, cond_scale=cond_scale)

    def synthesis_sample_text2semantic(self, grapheme_token_ids):
        return self.cfm_wrapper.sample(grapheme_token_ids=grapheme_token_ids)


@pytest.mark.parametrize("cfg", [syn.VOSINGLE, syn.VOMIX], ids=["vosingle", "vomix"])
def test_accelerate_acoustic_model_on_real_wrapper(ref_modules, monkeypatch, cfg):
    acoustic = ref_modules[0]
    net = acoustic.CoVoMix(dim=cfg.dim, dim_in=cfg.dim_in, num_phoneme_tokens=cfg.num_phoneme_tokens, depth=cfg.depth,
                           dim_head=cfg.dim_head, heads=cfg.heads, twocondition_oneoutput=cfg.twocondition_oneoutput)
    wrapper = acoustic.ConditionalFlowMatcherWrapper(CoVoMix=net, use_torchode=False, cond_drop_prob=0.3).eval()
    ref_sig = inspect.signature(wrapper.sample)
    monkeypatch.setattr(dropin, "B200FlowSampler", Recorder)
    model = dropin.accelerate_acoustic_model(Shim(wrapper))
    rec = wrapper._b200_sampler
    # configuration recovered from the live module == the configuration the reference was constructed with
    assert rec.cfg == cfg
    # solver settings come from the wrapper's own odeint_kwargs (acoustic.py:586-591): midpoint, h = 0.0625
    assert rec.kw == dict(torchdiffeq_ode_method="midpoint", ode_step_size=0.0625)
    assert rec.device == next(net.parameters()).device
    # the swapped entry point is reached through the unchanged CoVoMixModel.synthesis_sample forwarder
    tag, _, kw = model.synthesis_sample("ids", "cond", "mask", 0.7)
    assert tag == "b200" and kw == dict(phoneme_ids="ids", cond="cond", mask="mask", cond_scale=0.7)
    assert wrapper._reference_sample.__func__ is acoustic.ConditionalFlowMatcherWrapper.sample
    # the real sampler class accepts exactly the reference's keywords (plus the test-only y0)
    from covomix_b200.flow import B200FlowSampler
    ours = inspect.signature(B200FlowSampler.sample).parameters
    for name, prm in ref_sig.parameters.items():
        assert name in ours and ours[name].kind == prm.kind and ours[name].default == prm.default, name
    # the live weights pack into the library's blob: every tensor the C side binds is present
    blob = packing.pack_flow_weights(rec.sd, rec.cfg)
    assert bytes(blob[:8]) == b"COVOWTS1"
    names = packing.blob_entry_names(blob)
    for need in ("null_cond", "time.w", "emb.table", "embed.wx", "embed.wpc", "convpos.wT", "adaln.w", "final.gamma",
                 "pred.w", "L0.qkv.w", f"L{cfg.depth - 1}.skip.w", f"L{cfg.depth - 1}.ff2.b"):
        assert need in names, need


def test_accelerate_acoustic_model_custom_solver(ref_modules, monkeypatch):
    acoustic = ref_modules[0]
    cfg = syn.FlowConfig(dim=128, depth=2, heads=2)
    net = acoustic.CoVoMix(dim=cfg.dim, dim_in=cfg.dim_in, num_phoneme_tokens=cfg.num_phoneme_tokens, depth=cfg.depth,
                           dim_head=cfg.dim_head, heads=cfg.heads)
    wrapper = acoustic.ConditionalFlowMatcherWrapper(CoVoMix=net, use_torchode=False, torchdiffeq_ode_method="euler",
                                                     ode_step_size=0.03125)
    monkeypatch.setattr(dropin, "B200FlowSampler", Recorder)
    dropin.accelerate_acoustic_model(Shim(wrapper))
    assert wrapper._b200_sampler.cfg == cfg        # heads read from the live Attention module, not the default 16
    assert wrapper._b200_sampler.kw == dict(torchdiffeq_ode_method="euler", ode_step_size=0.03125)


def test_accelerate_generator_on_real_generator(ref_modules, monkeypatch):
    _, _, models, env = ref_modules
    with open(os.path.join(REF, "hifi-gan", "config_covomix.json")) as f:
        h = env.AttrDict(json.load(f))
    gen = models.Generator(h).eval()
    monkeypatch.setattr(dropin, "B200Generator", Recorder)
    for strip in (False, True):                      # the scripts call remove_weight_norm() before use; both layouts load
        if strip:
            gen.remove_weight_norm()
        rec = dropin.accelerate_generator(gen, h)
        assert rec.cfg == syn.HIFIGAN_COVOMIX and rec.kw == dict(h_format="fp16")
        blob = packing.pack_hifigan_weights(rec.sd, rec.cfg, "fp16")
        names = packing.blob_entry_names(blob)
        assert {"conv_pre.w", "ups.0.w", "ups.3.b", "rb.0.c1.0.w", "rb.11.c2.2.b", "conv_post.w"} <= set(names)
        if strip:
            # weight-norm folding (models.py:118-125) done by packing == the module's own remove_weight_norm()
            assert folded.keys() == rec.sd.keys()
            for k, v in rec.sd.items():
                assert torch.allclose(folded[k], v, rtol=1e-6, atol=1e-9), k
            assert np.mean(blob != first_blob) < 1e-3      # same blob up to last-bit rounding of a few fp16 weights
        else:
            folded = packing.fold_weight_norm(rec.sd)
        first_blob = blob
    # reference output shapes the host mirror reproduces: [B, 1, L] and, for an unbatched [80, T] mel, [1, L]
    with torch.inference_mode():
        assert tuple(gen(torch.zeros(80, 4)).shape) == (1, syn.HIFIGAN_COVOMIX.out_len(4))
        assert tuple(gen(torch.zeros(2, 80, 4)).shape) == (2, 1, syn.HIFIGAN_COVOMIX.out_len(4))


@pytest.mark.parametrize("cfg", [syn.COSINGLE, syn.COMIX], ids=["cosingle", "comix"])
def test_accelerate_text2semantic_on_real_wrapper(ref_modules, monkeypatch, cfg):
    t2s_mod = ref_modules[1]
    m = t2s_mod.TextToSemantic(dim=cfg.dim, source_depth=cfg.source_depth, target_depth=cfg.target_depth,
                               semantic_pad_id=-1, text_pad_id=0, heads=cfg.heads,
                               num_text_token_ids=cfg.num_text_token_ids,
                               num_semantic_token_ids=cfg.num_semantic_token_ids, no_source_transformer=False,
                               two_output=cfg.two_output, two_input=False,
                               target_transformer_dim=cfg.target_transformer_dim).eval()
    wrapper = t2s_mod.TextToSemanticWrapper(m)
    ref_sig = inspect.signature(wrapper.sample)
    import covomix_b200.t2s as t2s_host
    real = t2s_host.B200TextToSemantic
    monkeypatch.setattr(t2s_host, "B200TextToSemantic", Recorder)
    model = dropin.accelerate_text2semantic(Shim(wrapper))
    rec = wrapper._b200_t2s
    assert rec.cfg == cfg and rec.kw == dict(weight_format="bf16")
    tag, _, kw = model.synthesis_sample_text2semantic("text ids")
    assert tag == "b200" and kw == dict(grapheme_token_ids="text ids")
    ours = inspect.signature(real.sample).parameters
    for name, prm in ref_sig.parameters.items():
        assert name in ours and ours[name].default == prm.default, name
    blob = packing.pack_t2s_weights(rec.sd, rec.cfg, "bf16")
    assert bytes(blob[:8]) == b"COVOWTS1"
