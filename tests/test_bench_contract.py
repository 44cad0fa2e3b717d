"""bench.py's roofline inputs and workload table against the figures SURVEY.md section 8d states (CPU only)."""
import importlib.util
import json
import os

from conftest import ROOT

import covomix_b200  # noqa: F401
from covomix_b200 import synthetic as syn


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_flops_match_survey():
    b = _bench()
    # per token per pass F_tok(N) = 2 (L_lin + 16384 N): 242.1 MFLOP at N = 650 (VoSingle), 277.1 MFLOP at N = 1650
    # (VoMix); C2 = 10.07 TFLOP per utterance (32 NFE x 2 passes); vocoder 281.3 MFLOP per mel frame
    flow, voc = b.algorithmic_flops(syn.VOSINGLE, b.WORKLOADS["c2"])
    assert abs(flow / (650 * 64) / 242.1e6 - 1) < 5e-3
    assert abs(flow / 10.07e12 - 1) < 5e-3
    assert abs(voc / (500 * 281.3e6) - 1) < 1e-9
    flow3, voc3 = b.algorithmic_flops(syn.VOMIX, b.WORKLOADS["c3"])
    assert abs(flow3 / (8 * 1650 * 128) / 277.1e6 - 1) < 5e-3
    assert abs(flow3 / (8 * 58.5e12) - 1) < 5e-3             # 64 NFE: 58.5 TFLOP per item
    assert abs(voc3 / (8 * 1500 * 281.3e6) - 1) < 1e-9


def test_workloads_follow_baseline_configs():
    b = _bench()
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "VoSingle" in base["configs"][1] and "32 Euler" in base["configs"][1] and "10 s" in base["configs"][1]
    c2 = b.WORKLOADS["c2"]
    assert (c2["model"], c2["n_steps"], c2["B"], c2["N"] - c2["prompt"]) == ("vosingle", 32, 1, 500)      # 10 s at 50 Hz
    assert "VoMix" in base["configs"][2] and "64 Euler" in base["configs"][2] and "batch 8" in base["configs"][2]
    c3 = b.WORKLOADS["c3"]
    assert (c3["model"], c3["n_steps"], c3["B"], c3["N"] - c3["prompt"]) == ("vomix", 64, 8, 1500)        # 30 s at 50 Hz
    for name in ("c4", "c4p"):
        w = b.WORKLOADS[name]
        assert w["t2s"]["steps"] == w["N"] - w["prompt"] == 1500 and w["B"] == 8                           # 64 utterances / 8 GPUs
    assert b.WORKLOADS["c4p"]["t2s_sms"] > 0 and "t2s_sms" not in b.WORKLOADS["c4"]


def test_both_arms_report_the_same_config():
    """The driver compares the `config` objects of the B200 arm and the reference arm (`same_config`)."""
    b = _bench()
    for key, wl in b.WORKLOADS.items():
        t2s_sms = wl.get("t2s_sms", 0)
        for world in (1, 8):
            ref = b.config_for(wl, world, t2s_sms, None)                       # run_reference
            ours = b.config_for(wl, world, t2s_sms, 148 - t2s_sms if t2s_sms else None)   # bench_workload
            assert ref == ours and ref["workload"] == wl["name"] and ref["global_batch"] == wl["B"] * world
            assert ("t2s_assumption" in ref) == ("t2s" in wl)
