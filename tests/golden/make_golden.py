"""Generate tests/golden/*.npz by running the REAL reference modules (this container only).

    python tests/golden/make_golden.py

Imports ``covomix/covomix_model/acoustic.py`` and ``hifi-gan/models.py`` read-only from
/root/reference (stubs only for modules that are absent here and carry no arithmetic on this
path: ``matplotlib``, ``torchode``; ``torchdiffeq.odeint`` is third-party and absent -- the stub
routes to the oracle's restatement, so the ``*_sample`` vectors pin everything *except* the
solver, which has its own closed-form test).  Loads the seeded synthetic state dicts of
``neurips2024-covomix_b200/synthetic.py`` into the reference modules and stores outputs.
The GPU box has no /root/reference: only the .npz files travel.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import covomix_b200  # noqa: E402  (repo-root shim for the hyphenated package dir)
from covomix_b200 import synthetic as syn  # noqa: E402
from oracle import covomix_oracle as orc  # noqa: E402


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        sys.modules[name] = m
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
    to = types.ModuleType("torchode")
    to.Tsit5 = object
    sys.modules["torchode"] = to
    tde = types.ModuleType("torchdiffeq")

    def odeint(fn, y0, t, *, atol=None, rtol=None, method="midpoint", options=None):
        return orc.odeint_fixed_grid(fn, y0, t, method=method, step_size=options["step_size"])

    tde.odeint = odeint
    sys.modules["torchdiffeq"] = tde


def ref_flow(cfg: syn.FlowConfig, sd):
    sys.path.insert(0, REF)
    from covomix.covomix_model.acoustic import CoVoMix, ConditionalFlowMatcherWrapper
    net = CoVoMix(dim=cfg.dim, dim_in=cfg.dim_in, num_phoneme_tokens=cfg.num_phoneme_tokens, depth=cfg.depth,
                  dim_head=cfg.dim_head, heads=cfg.heads, twocondition_oneoutput=cfg.twocondition_oneoutput)
    net.load_state_dict(sd, strict=True)
    wrapper = ConditionalFlowMatcherWrapper(CoVoMix=net, use_torchode=False, cond_drop_prob=0.3)
    return wrapper.eval()


def ref_generator(cfg: syn.HifiganConfig, sd):
    sys.path.insert(0, os.path.join(REF, "hifi-gan"))
    from env import AttrDict
    from models import Generator
    with open(os.path.join(REF, "hifi-gan", "config_covomix.json")) as f:
        h = AttrDict(json.load(f))
    gen = Generator(h).eval()
    gen.remove_weight_norm()
    gen.load_state_dict(sd, strict=True)
    return gen


@torch.inference_mode()
def main():
    _install_stubs()
    torch.manual_seed(0)
    out = {}

    for name, cfg, B, N in (("vosingle", syn.VOSINGLE, 1, 96), ("vomix", syn.VOMIX, 2, 80)):
        sd = syn.synthetic_flow_state_dict(cfg, seed=1234)
        wrapper = ref_flow(cfg, sd)
        ids, cond, y0, mask = syn.synthetic_flow_inputs(cfg, B, N, prompt=24, seed=30)
        t = torch.tensor(0.28125)
        v = wrapper.CoVoMix.forward_with_cond_scale(y0, times=t, phoneme_ids=ids, cond=cond, cond_scale=0.7)
        v_cond = wrapper.CoVoMix.forward(y0, times=t, phoneme_ids=ids, cond=cond, cond_drop_prob=0.)
        # sample(): y0 comes from torch.randn_like inside the reference; replay the same draw.
        torch.manual_seed(77)
        y0_s = torch.randn_like(cond if not cfg.twocondition_oneoutput else cond[:, :, :80])
        torch.manual_seed(77)
        mel = wrapper.sample(phoneme_ids=ids, cond=cond, mask=mask, cond_scale=0.7)
        np.savez_compressed(os.path.join(HERE, f"flow_{name}.npz"),
                            B=B, N=N, prompt=24, weight_seed=1234, input_seed=30, t=0.28125, cond_scale=0.7,
                            v_cfg=v.numpy(), v_cond=v_cond.numpy(), y0_sample=y0_s.numpy(), mel=mel.numpy())
        print(name, "v", v.shape, float(v.std()), "mel", mel.shape, float(mel.std()))
        del wrapper, sd

    # multi-tile cases (round 2): B = 2, N = 300 -> three 128-key tiles and two query tiles per (sequence, head) inside
    # the network, so the attention tiling is pinned to the reference module (acoustic.py:430-521), not only to the oracle
    for name, cfg, B, N in (("vosingle_n300", syn.VOSINGLE, 2, 300), ("vomix_n300", syn.VOMIX, 2, 300)):
        sd = syn.synthetic_flow_state_dict(cfg, seed=1234)
        wrapper = ref_flow(cfg, sd)
        ids, cond, y0, mask = syn.synthetic_flow_inputs(cfg, B, N, prompt=60, seed=31)
        t = torch.tensor(0.59375)
        v = wrapper.CoVoMix.forward_with_cond_scale(y0, times=t, phoneme_ids=ids, cond=cond, cond_scale=0.7)
        torch.manual_seed(78)
        y0_s = torch.randn_like(cond if not cfg.twocondition_oneoutput else cond[:, :, :80])
        torch.manual_seed(78)
        mel = wrapper.sample(phoneme_ids=ids, cond=cond, mask=mask, cond_scale=0.7)
        # y0 of sample() is the torch.manual_seed(78) CPU draw; tests replay it instead of storing it
        np.savez_compressed(os.path.join(HERE, f"flow_{name}.npz"),
                            B=B, N=N, prompt=60, weight_seed=1234, input_seed=31, t=0.59375, cond_scale=0.7, y0_seed=78,
                            v_cfg=v.numpy(), mel=mel.numpy())
        print(name, "v", v.shape, float(v.std()), "mel", mel.shape, float(mel.std()))
        del wrapper, sd

    hcfg = syn.HIFIGAN_COVOMIX
    sd = syn.synthetic_hifigan_state_dict(hcfg, seed=1234)
    gen = ref_generator(hcfg, sd)
    g = torch.Generator().manual_seed(30)
    mel_c1 = syn.synthetic_logmel(g, 1, 80, 256)
    mel_u = syn.synthetic_logmel(g, 80, 64)
    mel_b = syn.synthetic_logmel(g, 2, 80, 48)
    w1, wu, wb = gen(mel_c1), gen(mel_u), gen(mel_b)
    print("hifigan", w1.shape, float(w1.std()), wu.shape, wb.shape)
    np.savez_compressed(os.path.join(HERE, "hifigan.npz"), weight_seed=1234, input_seed=30,
                        wav_c1=w1.numpy(), wav_unbatched=wu.numpy(), wav_batch=wb.numpy())


if __name__ == "__main__":
    main()
