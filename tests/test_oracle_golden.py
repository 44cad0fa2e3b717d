"""The oracle (oracle/covomix_oracle.py) against vectors produced by the real reference
modules (tests/golden/make_golden.py), plus the closed-form test of the ODE restatement."""
import os

import numpy as np
import pytest
import torch

import covomix_b200  # noqa: F401
from covomix_b200 import synthetic as syn
from oracle import covomix_oracle as orc
from conftest import GOLDEN


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name,cfg", [("vosingle", syn.VOSINGLE), ("vomix", syn.VOMIX)])
def test_velocity_matches_reference(name, cfg):
    g = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    sd = syn.synthetic_flow_state_dict(cfg, seed=int(g["weight_seed"]))
    ids, cond, y0, _ = syn.synthetic_flow_inputs(cfg, int(g["B"]), int(g["N"]), prompt=int(g["prompt"]),
                                                 seed=int(g["input_seed"]))
    t = torch.tensor(float(g["t"]))
    with torch.inference_mode():
        v_cond = orc.velocity(sd, cfg, y0, ids, cond, t, drop_cond=False)
        v_cfg = orc.velocity_cfg(sd, cfg, y0, ids, cond, t, float(g["cond_scale"]))
    # same fp32 ops, possibly different summation order: tolerance 1e-5 relative
    assert rel_l2(v_cond, g["v_cond"]) < 1e-5
    assert rel_l2(v_cfg, g["v_cfg"]) < 1e-5


@pytest.mark.parametrize("name,cfg", [("vosingle_n300", syn.VOSINGLE), ("vomix_n300", syn.VOMIX)])
def test_velocity_matches_reference_multi_tile(name, cfg):
    """B = 2, N = 300 vectors from the real reference module (three key tiles / two query tiles for the CUDA kernel)."""
    g = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    sd = syn.synthetic_flow_state_dict(cfg, seed=int(g["weight_seed"]))
    ids, cond, y0, _ = syn.synthetic_flow_inputs(cfg, int(g["B"]), int(g["N"]), prompt=int(g["prompt"]),
                                                 seed=int(g["input_seed"]))
    v = orc.velocity_cfg(sd, cfg, y0, ids, cond, torch.tensor(float(g["t"])), float(g["cond_scale"]))
    assert rel_l2(v, g["v_cfg"]) < 1e-5


@pytest.mark.parametrize("name,cfg", [("vosingle", syn.VOSINGLE), ("vomix", syn.VOMIX)])
def test_sample_matches_reference(name, cfg):
    g = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    sd = syn.synthetic_flow_state_dict(cfg, seed=int(g["weight_seed"]))
    ids, cond, _, _ = syn.synthetic_flow_inputs(cfg, int(g["B"]), int(g["N"]), prompt=int(g["prompt"]),
                                                seed=int(g["input_seed"]))
    mel = orc.flow_sample(sd, cfg, ids, cond, torch.from_numpy(g["y0_sample"]), cond_scale=float(g["cond_scale"]))
    assert rel_l2(mel, g["mel"]) < 1e-4


def test_hifigan_matches_reference():
    g = np.load(os.path.join(GOLDEN, "hifigan.npz"))
    cfg = syn.HIFIGAN_COVOMIX
    sd = syn.synthetic_hifigan_state_dict(cfg, seed=int(g["weight_seed"]))
    gen = torch.Generator().manual_seed(int(g["input_seed"]))
    mel_c1 = syn.synthetic_logmel(gen, 1, 80, 256)
    mel_u = syn.synthetic_logmel(gen, 80, 64)
    mel_b = syn.synthetic_logmel(gen, 2, 80, 48)
    for mel, key in ((mel_c1, "wav_c1"), (mel_u, "wav_unbatched"), (mel_b, "wav_batch")):
        wav = orc.hifigan_forward(sd, cfg, mel)
        ref = g[key]
        assert tuple(wav.shape) == ref.shape        # [1, L] for the unbatched [80, T] input, [B, 1, L] otherwise
        assert wav.shape[-1] == cfg.out_len(mel.shape[-1]) == 160 * mel.shape[-1] + 32
        assert rel_l2(wav.reshape(-1), ref.reshape(-1)) < 1e-5


def test_odeint_known_answer():
    """dy/dt = a*y: 16 midpoint steps give y0*(1 + a/16 + a^2/512)^16; 32 Euler steps y0*(1+a/32)^32."""
    a = -1.7
    y0 = torch.tensor([1.0, -2.0, 0.5])
    fn = lambda t, y: a * y
    t = torch.linspace(0, 1, 3)
    sol = orc.odeint_fixed_grid(fn, y0, t, method="midpoint", step_size=0.0625)
    assert sol.shape[0] == 3
    assert torch.allclose(sol[-1], y0 * (1 + a / 16 + a * a / 512) ** 16, rtol=1e-5)
    assert torch.allclose(sol[1], y0 * (1 + a / 16 + a * a / 512) ** 8, rtol=1e-5)
    sol = orc.odeint_fixed_grid(fn, y0, t, method="euler", step_size=1 / 32)
    assert torch.allclose(sol[-1], y0 * (1 + a / 32) ** 32, rtol=1e-5)


def test_odeint_nfe_and_times():
    seen = []

    def fn(t, y):
        seen.append(float(t))
        return torch.zeros_like(y)

    orc.odeint_fixed_grid(fn, torch.zeros(2), torch.linspace(0, 1, 3))
    assert len(seen) == 32
    assert seen == pytest.approx([k / 32 for k in range(32)])
    assert list(orc.ode_eval_times("midpoint", 16)) == pytest.approx(seen)
