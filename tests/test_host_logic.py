"""CPU tests of the host side: packed-weight layout, implicit-GEMM index maps (emulated in torch),
weight-norm folding, config inference, C-ABI surface, drop-in swap, sharding."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import covomix_b200  # noqa: F401
from covomix_b200 import _native as nat, packing, sharding, synthetic as syn
from conftest import ROOT


def parse_blob(blob: np.ndarray):
    assert bytes(blob[:8]) == b"COVOWTS1"
    n = struct.unpack("<I", bytes(blob[8:12]))[0]
    out = {}
    for i in range(n):
        e = packing._ENTRY.unpack(bytes(blob[16 + i * packing._ENTRY.size:16 + (i + 1) * packing._ENTRY.size]))
        name = e[0].split(b"\0")[0].decode()
        dtype, ndim, shape, off, nbytes = e[1], e[2], e[3:7], e[7], e[8]
        raw = blob[off:off + nbytes]
        shape = tuple(int(s) for s in shape[:ndim])
        if dtype == packing.DT_F32:
            t = torch.from_numpy(raw.view(np.float32).copy()).reshape(shape)
        elif dtype == packing.DT_BF16:
            t = torch.from_numpy(raw.view(np.int16).copy()).view(torch.bfloat16).reshape(shape).float()
        else:
            t = torch.from_numpy(raw.view(np.int16).copy()).view(torch.float16).reshape(shape).float()
        assert off % 256 == 0
        out[name] = t
    return out


SMALL_HIFI = syn.HifiganConfig(upsample_rates=(5, 4, 2), upsample_kernel_sizes=(8, 8, 4), upsample_initial_channel=24,
                               resblock_kernel_sizes=(3, 7), resblock_dilation_sizes=((1, 3), (1, 3)), num_mels=10)


def emulate_hifigan_from_blob(w, cfg, mel):
    """Executes Generator.forward with exactly the index maps of csrc/hifigan.cuh + gemm_sm100.cuh
    (time-major activations, tap-major packed weights, polyphase transposed conv) in fp32 torch."""
    pad = lambda c: (c + 63) // 64 * 64
    B, C, T = mel.shape

    def conv(act, wp, bias, k, dil):                       # act [B,T,Cin_pad]; wp [Cout_pad, k*Cin_pad]
        Bq, Tq, cin = act.shape
        p = (k * dil - dil) // 2
        out = bias.expand(Bq, Tq, -1).clone()
        for j in range(k):
            off = j * dil - p
            sh = torch.zeros_like(act)
            lo, hi = max(0, -off), min(Tq, Tq - off)
            if hi > lo:
                sh[:, lo:hi] = act[:, lo + off:hi + off]
            out += sh @ wp[:, j * cin:(j + 1) * cin].t()
        return out

    x = torch.zeros(B, T, pad(C))
    x[:, :, :C] = mel.transpose(1, 2)
    a = F.leaky_relu(conv(x, w["conv_pre.w"], w["conv_pre.b"], 7, 1), 0.1)
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        wp, bp = w[f"ups.{i}.w"], w[f"ups.{i}.b"]
        J, p = (k + u - 1) // u, (k - u) // 2
        cin = a.shape[-1]
        cop = wp.shape[0] // u
        t_in = a.shape[1]
        t_out = (t_in - 1) * u - 2 * p + k
        xs = torch.zeros(B, t_out, cop)
        if k - 2 * p == u:
            # aligned polyphase form (hifigan.cuh: upconv_align_weights_kernel + the plan's tap rows): weights re-indexed
            # from the blob's matrix as r = (rr + p) % u, j = m + (rr + p) // u; GEMM row q = output samples u*q .. u*q + u - 1
            m0 = -1 if p > 0 else 0
            taps = max((k - 1 - rr - p) // u for rr in range(u)) - m0 + 1
            al = torch.zeros(u, cop, taps, cin)
            pw = wp.reshape(u, cop, J, cin)
            for rr in range(u):
                for mt in range(taps):
                    j = m0 + mt + (rr + p) // u
                    if 0 <= j < J:
                        al[rr, :, mt] = pw[(rr + p) % u, :, j]
            al = al.reshape(u * cop, taps * cin)
            assert t_out == u * t_in
            for q in range(t_in):
                d = bp.expand(B, -1).clone()
                for mt in range(taps):
                    row = q - (m0 + mt)                              # tap_row = -(m0 + mt); rows outside [0, t_in) are zero-filled
                    if 0 <= row < t_in:
                        d += a[:, row] @ al[:, mt * cin:(mt + 1) * cin].t()
                xs[:, u * q:u * q + u] = d.reshape(B, u, cop)
        for q in range(t_in + J - 1 if k - 2 * p != u else 0):
            d = bp.expand(B, -1).clone()
            for j in range(J):
                if 0 <= q - j < t_in:
                    d += a[:, q - j] @ wp[:, j * cin:(j + 1) * cin].t()
            for r in range(u):
                o = u * q + r - p
                if 0 <= o < t_out:
                    xs[:, o] = d[:, r * cop:(r + 1) * cop]
        x = xs
        a0 = F.leaky_relu(x, 0.1)
        acc = None
        for j in range(nk):
            kk = cfg.resblock_kernel_sizes[j]
            xr, ar = x, a0
            for m, dil in enumerate(cfg.resblock_dilation_sizes[j]):
                r = i * nk + j
                h = F.leaky_relu(conv(ar, w[f"rb.{r}.c1.{m}.w"], w[f"rb.{r}.c1.{m}.b"], kk, dil), 0.1)
                xr = conv(h, w[f"rb.{r}.c2.{m}.w"], w[f"rb.{r}.c2.{m}.b"], kk, 1) + xr
                ar = F.leaky_relu(xr, 0.1)
            acc = xr if acc is None else acc + xr
        a = F.leaky_relu(acc / nk, 0.01 if i + 1 == len(cfg.upsample_rates) else 0.1)
    wpost, bpost = w["conv_post.w"], w["conv_post.b"]
    Tq = a.shape[1]
    y = torch.full((B, Tq), float(bpost[0]))
    for kx in range(7):
        off = kx - 3
        lo, hi = max(0, -off), min(Tq, Tq - off)
        y[:, lo:hi] += a[:, lo + off:hi + off] @ wpost[kx]
    return torch.tanh(y)[:, None, :]


def test_hifigan_packing_and_index_maps_match_oracle():
    from oracle import covomix_oracle as orc
    cfg = SMALL_HIFI
    sd = syn.synthetic_hifigan_state_dict(cfg, 7)
    blob = packing.pack_hifigan_weights(sd, cfg, "fp16")
    w = parse_blob(blob)
    # fp16 storage rounds the weights; compare against the oracle run on the same rounded weights
    sd16 = {k: (v.half().float() if k.endswith(".weight") and not k.startswith("conv_post") else v) for k, v in sd.items()}
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(3), 2, cfg.num_mels, 9)
    got = emulate_hifigan_from_blob(w, cfg, mel)
    ref = orc.hifigan_forward(sd16, cfg, mel)
    assert got.shape == ref.shape == (2, 1, cfg.out_len(9))
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4)


def test_fold_weight_norm_matches_torch():
    torch.manual_seed(0)
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(6, 4, 3))
    convt = torch.nn.utils.weight_norm(torch.nn.ConvTranspose1d(6, 4, 4, 2))
    with torch.no_grad():
        conv.weight_g.mul_(1.7)
        convt.weight_g.mul_(0.3)
    sd = {"a." + k: v for k, v in conv.state_dict().items()}
    sd.update({"b." + k: v for k, v in convt.state_dict().items()})
    folded = packing.fold_weight_norm(sd)
    torch.nn.utils.remove_weight_norm(conv)
    torch.nn.utils.remove_weight_norm(convt)
    assert torch.allclose(folded["a.weight"], conv.weight, atol=1e-6)
    assert torch.allclose(folded["b.weight"], convt.weight, atol=1e-6)
    assert set(folded) == {"a.weight", "a.bias", "b.weight", "b.bias"}


SMALL_FLOW = syn.FlowConfig(dim=128, depth=4, heads=2, dim_in=160, twocondition_oneoutput=True, dim_phoneme_emb=64)


def test_flow_packing_layout():
    cfg = SMALL_FLOW
    sd = syn.synthetic_flow_state_dict(cfg, 5)
    lightning = {"cfm_wrapper.CoVoMix." + k: v for k, v in sd.items()}
    w = parse_blob(packing.pack_flow_weights(lightning, cfg))
    W = sd["to_embed.weight"]
    assert w["embed.wx"].shape == (128, 128) and w["embed.wpc"].shape == (128, 320)
    assert torch.equal(w["embed.wx"][:, :80], W[:, :80].bfloat16().float()) and w["embed.wx"][:, 80:].abs().sum() == 0
    assert torch.equal(w["embed.wpc"][:, :288], W[:, 80:].bfloat16().float())
    # AdaLN stack order per layer: gamma1 | beta1 | gamma2 | beta2
    D = cfg.dim
    assert torch.equal(w["adaln.w"][5 * D:6 * D], sd["transformer.layers.1.1.to_beta.weight"])
    assert torch.equal(w["adaln.b"][6 * D:7 * D], sd["transformer.layers.1.3.to_gamma.bias"])
    assert torch.equal(w["convpos.wT"][3], sd["conv_embed.dw_conv1d.0.weight"][:, 0, 3])
    assert "L1.skip.w" not in w and w["L2.skip.w"].shape == (D, 2 * D)
    assert w["pred.w"].shape == (128, D) and w["pred.w"][80:].abs().sum() == 0
    assert packing.flow_config_from_state_dict(lightning, heads=2) == cfg
    assert packing.flow_config_from_state_dict(syn.synthetic_flow_state_dict(syn.FlowConfig(dim=128, depth=2, heads=2,
                                               dim_phoneme_emb=64), 1), heads=2).twocondition_oneoutput is False


def test_abi_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "covomix_b200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(covo_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    lib = nat.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.covo_version() == 1
    assert ctypes.sizeof(nat.FlowCfg) == 44 and ctypes.sizeof(nat.HifiganCfg) == 4 * (3 + 8 + 8 + 1 + 4 + 1 + 16 + 2)


def test_native_has_no_cpu_fallback():
    from covomix_b200.flow import B200FlowSampler
    from covomix_b200.vocoder import B200Generator
    with pytest.raises(RuntimeError):
        B200FlowSampler({}, syn.VOSINGLE, device="cpu")
    with pytest.raises(RuntimeError):
        B200Generator({}, syn.HIFIGAN_COVOMIX, device="cpu")
    if not torch.cuda.is_available():
        # handle creation on a machine without a GPU must fail loudly, not fall back
        sd = syn.synthetic_hifigan_state_dict(SMALL_HIFI, 1)
        with pytest.raises(RuntimeError):
            B200Generator(sd, SMALL_HIFI, device="cuda:0")


def test_dropin_swaps_sample(monkeypatch):
    from covomix_b200 import dropin

    class FakeSampler:
        def __init__(self, sd, cfg, device, **kw):
            self.cfg, self.kw = cfg, kw

        def sample(self, *, phoneme_ids, cond, mask=None, steps=3, cond_scale=1., decode_to_audio=False):
            return "b200"

    monkeypatch.setattr(dropin, "B200FlowSampler", FakeSampler)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for k, v in syn.synthetic_flow_state_dict(SMALL_FLOW, 1).items():
                self.register_buffer(k.replace(".", "__"), v)

        def state_dict(self, *a, **k):
            return {n.replace("__", "."): v for n, v in super().state_dict().items()}

        def parameters(self, recurse=True):
            return iter([torch.nn.Parameter(torch.zeros(1))])

    class Wrapper:
        def __init__(self):
            self.CoVoMix = Net()
            self.odeint_kwargs = dict(atol=1e-5, rtol=1e-5, method="midpoint", options=dict(step_size=0.0625))

        def sample(self, **kw):
            return "reference"

    class Model:
        def __init__(self):
            self.cfm_wrapper = Wrapper()

        def synthesis_sample(self, phoneme_ids, cond, mask, cond_scale):      # conditional_model.py:295-302
            return self.cfm_wrapper.sample(phoneme_ids=phoneme_ids, cond=cond, mask=mask, cond_scale=cond_scale)

    m = Model()
    assert m.synthesis_sample(None, None, None, 0.7) == "reference"
    dropin.accelerate_acoustic_model(m, device="cuda:0", heads=2)
    assert m.synthesis_sample(None, None, None, 0.7) == "b200"
    s = m.cfm_wrapper._b200_sampler
    assert s.cfg == SMALL_FLOW and s.kw == dict(torchdiffeq_ode_method="midpoint", ode_step_size=0.0625)


def test_sharding_plan():
    lengths = [1650] * 10 + [650] * 5 + [900]
    per_rank = sharding.assign_batches(lengths, world=4, batch=4)
    seen = sorted(i for r in per_rank for _, idx in r for i in idx)
    assert seen == list(range(len(lengths)))
    for r in per_rank:
        for n, idx in r:
            assert 1 <= len(idx) <= 4 and all(lengths[i] == n for i in idx)
    loads = [sum(sharding.batch_cost(len(idx), n) for n, idx in r) for r in per_rank]
    assert max(loads) <= 2.0 * (sum(loads) / 4)
    assert sharding.assign_batches([], 2, 8) == [[], []]
