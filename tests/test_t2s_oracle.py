"""The text-to-semantic oracle (oracle/t2s_oracle.py) against vectors produced by the real reference
``TextToSemantic.generate`` (tests/golden/make_golden_t2s.py)."""
import os

import numpy as np
import pytest
import torch

import covomix_b200  # noqa: F401
from covomix_b200 import synthetic as syn
from oracle import t2s_oracle as t2s
from conftest import GOLDEN

CASES = [("comix_b2", syn.COMIX), ("comix_eos", syn.COMIX), ("cosingle_b1", syn.COSINGLE)]


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def load_case(name, cfg):
    g = np.load(os.path.join(GOLDEN, f"t2s_{name}.npz"))
    sd = syn.synthetic_t2s_state_dict(cfg, seed=int(g["weight_seed"]))
    ids = syn.synthetic_text_ids(cfg, int(g["B"]), int(g["S"]), seed=int(g["input_seed"]), ragged=bool(g["ragged"]))
    return g, sd, ids


@pytest.mark.parametrize("name,cfg", CASES)
def test_generate_matches_reference(name, cfg):
    g, sd, ids = load_case(name, cfg)
    with torch.inference_mode():
        enc, mask = t2s.encode(sd, cfg, ids)
    assert rel_l2(enc, g["enc"]) < 1e-5
    target, tmask, steps, logits = t2s.generate(sd, cfg, ids, torch.from_numpy(g["u"]), max_length=int(g["max_length"]),
                                                return_logits=True)
    assert steps == int(g["steps"])
    # same fp32 ops, possibly different summation order (cached cross k/v, per-position decode)
    assert rel_l2(logits, g["logits"]) < 1e-5
    assert torch.equal(target, torch.from_numpy(g["target"]))
    assert torch.equal(tmask, torch.from_numpy(g["mask"]))


def test_eos_case_stops_early_and_keeps_eos():
    g, sd, ids = load_case("comix_eos", syn.COMIX)
    assert int(g["steps"]) < int(g["max_length"])
    out = t2s.sample(sd, syn.COMIX, ids, torch.from_numpy(g["u"]), max_length=int(g["max_length"]))
    assert int((out == syn.COMIX.semantic_eos_id).sum()) == 1          # the EOS token itself is returned
    assert out.numel() == int(g["mask"].sum())


def test_teacher_forcing_reproduces_free_running():
    g, sd, ids = load_case("cosingle_b1", syn.COSINGLE)
    steps = int(g["steps"])
    forced = torch.from_numpy(g["target"]).view(1, 1, steps)
    u = torch.full((steps, 1, 1, syn.COSINGLE.n_logits), 0.5)           # different noise: sampled tokens differ, logits do not
    _, _, _, logits = t2s.generate(sd, syn.COSINGLE, ids, u, max_length=steps, forced=forced, return_logits=True)
    assert rel_l2(logits, g["logits"]) < 1e-5


def test_rotary_is_interleaved_not_neox():
    x = torch.arange(8.0).view(1, 1, 1, 8)
    inv = torch.tensor([1.0, 0.5, 0.25, 0.125])
    y = t2s.rotary(x, inv, torch.tensor([1]))
    c, s = torch.cos(torch.tensor(1.0)), torch.sin(torch.tensor(1.0))
    assert torch.allclose(y[0, 0, 0, 0], 0 * c - 1 * s) and torch.allclose(y[0, 0, 0, 1], 1 * c + 0 * s)
