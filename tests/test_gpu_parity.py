"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors produced by the
real reference modules and against the CPU oracle on the same seeded inputs.

Tolerances (reference arithmetic is fp32; ours is bf16/fp16 operands with fp32 accumulation):
  one velocity evaluation   rel-L2 <= 1e-2, max-abs <= 5e-2 * std(v)
  sampled mel (32 NFE)      rel-L2 <= 3e-2, mean-abs <= 0.05
  waveform                  rel-L2 <= 2e-3 (fp16 operands, default) / 1e-2 (bf16 operands)
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import covomix_b200  # noqa: F401
from covomix_b200 import _native as nat, synthetic as syn
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)


def record(what, **numbers):
    """Measured parity numbers -> stdout and gpurun_out/parity_numbers.jsonl (quoted in DESIGN.md section 2)."""
    import json
    line = {"what": what, **{k: float(f"{v:.4g}") for k, v in numbers.items()}}
    print("PARITY", json.dumps(line))
    out = os.path.join(os.path.dirname(GOLDEN), os.pardir, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_numbers.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    except OSError:
        pass


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def orc():
    from oracle import covomix_oracle
    return covomix_oracle


@pytest.fixture(scope="module")
def vosingle(dev):
    from covomix_b200.flow import B200FlowSampler
    sd = syn.synthetic_flow_state_dict(syn.VOSINGLE, 1234)
    return sd, B200FlowSampler(sd, syn.VOSINGLE, dev)


@pytest.fixture(scope="module")
def vomix(dev):
    from covomix_b200.flow import B200FlowSampler
    sd = syn.synthetic_flow_state_dict(syn.VOMIX, 1234)
    return sd, B200FlowSampler(sd, syn.VOMIX, dev)


@pytest.fixture(scope="module")
def vocoder(dev):
    from covomix_b200.vocoder import B200Generator
    sd = syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234)
    return sd, B200Generator(sd, syn.HIFIGAN_COVOMIX, dev)


# ------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (1, 64, 64, 0), (129, 256, 128, 128), (300, 512, 1024, 0),
                                      (2600, 3072, 1024, 256), (1300, 1024, 4096, 0)])
def test_gemm_kernel(dev, M, N, K, bn):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    bias, res = torch.randn(N, device=dev), torch.randn(M, N, device=dev)
    out = torch.full((M, N), float("nan"), device=dev)
    outh = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    nat.check(nat.lib().covo_dbg_gemm(P(A), P(W), P(bias), P(res), P(out), P(outh), M, N, K, 1, bn, None), "gemm")
    ref = A.float() @ W.float().t() + bias + res
    assert rel_l2(out, ref) < 1e-5                      # same bf16 inputs, fp32 accumulate: only summation order differs
    assert rel_l2(outh.float(), torch.nn.functional.gelu(ref)) < 4e-3


@pytest.mark.parametrize("Bt,N,H", [(1, 1, 1), (1, 127, 2), (2, 128, 1), (1, 129, 1), (3, 650, 16), (2, 1650, 16)])
def test_attention_kernel(dev, Bt, N, H):
    torch.manual_seed(N)
    qkv = torch.randn(Bt, N, 3 * H * 64, device=dev).bfloat16()
    out = torch.zeros(Bt, N, H * 64, device=dev, dtype=torch.bfloat16)
    nat.check(nat.lib().covo_dbg_attention(P(qkv), P(out), Bt, N, H, 0, None), "attention")
    q, k, v = (t.reshape(Bt, N, H, 64).permute(0, 2, 1, 3).float() for t in qkv.chunk(3, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(Bt, N, H * 64)
    assert rel_l2(out.float(), ref) < 5e-3              # P and O rounded to bf16


# ------------------------------------------------------------------------------------------ flow vs reference golden
@pytest.mark.parametrize("name", ["vosingle", "vomix"])
def test_velocity_matches_reference(dev, name, vosingle, vomix):
    sd, smp = vosingle if name == "vosingle" else vomix
    g = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    ids, cond, y0, _ = syn.synthetic_flow_inputs(smp.cfg, int(g["B"]), int(g["N"]), prompt=int(g["prompt"]),
                                                 seed=int(g["input_seed"]))
    v = smp.velocity(y0.to(dev), times=float(g["t"]), phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    ref = torch.from_numpy(g["v_cfg"])
    assert rel_l2(v, ref) < 1e-2
    assert float((v.cpu() - ref).abs().max()) < 5e-2 * float(ref.std())
    v1 = smp.velocity(y0.to(dev), times=float(g["t"]), phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=1.0)
    assert rel_l2(v1, g["v_cond"]) < 1e-2               # cond_scale == 1 -> conditional branch only (acoustic.py:423)


@pytest.mark.parametrize("name", ["vosingle", "vomix"])
def test_sample_matches_reference(dev, name, vosingle, vomix):
    sd, smp = vosingle if name == "vosingle" else vomix
    g = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    ids, cond, _, mask = syn.synthetic_flow_inputs(smp.cfg, int(g["B"]), int(g["N"]), prompt=int(g["prompt"]),
                                                   seed=int(g["input_seed"]))
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), mask=mask.to(dev), cond_scale=0.7,
                     y0=torch.from_numpy(g["y0_sample"]).to(dev))       # reference default: midpoint, h = 1/16, 32 NFE
    ref = torch.from_numpy(g["mel"])
    assert mel.shape == ref.shape
    assert rel_l2(mel, ref) < 3e-2
    assert float((mel.cpu() - ref).abs().mean()) < 0.05
    # calling again (CUDA-graph replay) is deterministic
    mel2 = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), mask=mask.to(dev), cond_scale=0.7,
                      y0=torch.from_numpy(g["y0_sample"]).to(dev))
    assert torch.equal(mel, mel2)


# ------------------------------------------------------------------------------------------ multi-tile goldens (N = 300)
@pytest.mark.parametrize("name", ["vosingle_n300", "vomix_n300"])
def test_multi_tile_matches_reference(dev, name, vosingle, vomix):
    """B = 2, N = 300 vectors of the REAL reference module (tests/golden/make_golden.py): three 128-key tiles and two
    query tiles per (sequence, head) inside the network -- velocity and the sampled mel (midpoint, 32 NFE)."""
    sd, smp = vosingle if name.startswith("vosingle") else vomix
    g = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    ids, cond, y0, mask = syn.synthetic_flow_inputs(smp.cfg, int(g["B"]), int(g["N"]), prompt=int(g["prompt"]),
                                                    seed=int(g["input_seed"]))
    v = smp.velocity(y0.to(dev), times=float(g["t"]), phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    ref = torch.from_numpy(g["v_cfg"])
    r, mx = rel_l2(v, ref), float((v.cpu() - ref).abs().max()) / float(ref.std())
    record(f"velocity {name} (reference golden)", rel_l2=r, max_abs_over_sigma=mx)
    assert r < 1e-2 and mx < 5e-2
    torch.manual_seed(int(g["y0_seed"]))
    y0_s = torch.randn_like(cond if smp.cfg.n_streams == 1 else cond[:, :, :80])       # the draw sample() made in the reference
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), mask=mask.to(dev), cond_scale=0.7, y0=y0_s.to(dev))
    ref = torch.from_numpy(g["mel"])
    r, ma = rel_l2(mel, ref), float((mel.cpu() - ref).abs().mean())
    record(f"sample {name}, midpoint 32 NFE (reference golden)", rel_l2=r, mean_abs=ma)
    assert r < 3e-2 and ma < 0.05


# ------------------------------------------------------------------------------------------ BASELINE shapes vs the oracle
@pytest.mark.parametrize("name,N", [("vosingle", 650), ("vomix", 1650)])
def test_velocity_at_baseline_shapes_vs_oracle(dev, orc, name, N, vosingle, vomix):
    """One CFG velocity evaluation at the C2 shape (VoSingle, B = 1, N = 650) and at C3's per-item shape (VoMix, B = 1,
    N = 1650: 13 key tiles, 7 query tiles) against the fp32 oracle run in-test.  SURVEY 8d tolerances."""
    sd, smp = vosingle if name == "vosingle" else vomix
    ids, cond, y0, _ = syn.synthetic_flow_inputs(smp.cfg, 1, N, prompt=150, seed=30)
    v = smp.velocity(y0.to(dev), times=0.40625, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    with torch.inference_mode():
        ref = orc.velocity_cfg(sd, smp.cfg, y0, ids, cond, torch.tensor(0.40625), 0.7)
    r, mx = rel_l2(v, ref), float((v.cpu() - ref).abs().max()) / float(ref.std())
    # "tolerances to be confirmed against the measured bf16-PyTorch-on-GPU error of the same model": the oracle
    # itself on the GPU under torch.autocast(bfloat16) (plain torch ops; test-side only) against its fp32 CPU result
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        sdg = {k: t.to(dev) for k, t in sd.items()}
        v16 = orc.velocity_cfg(sdg, smp.cfg, y0.to(dev), ids.to(dev), cond.to(dev), torch.tensor(0.40625, device=dev), 0.7).float()
    r16 = rel_l2(v16, ref)
    record(f"velocity {name} B=1 N={N} (oracle)", rel_l2=r, max_abs_over_sigma=mx, torch_bf16_on_gpu_rel_l2=r16)
    assert r < 1e-2 and mx < 5e-2
    assert r < 1.5 * r16    # same error class as PyTorch's own bf16 autocast run of the model


def test_c2_full_sample_vs_oracle(dev, orc, vosingle):
    """BASELINE configs[1] end to end: VoSingle, 32 Euler steps, N = 650 (10 s + 3 s prompt), B = 1, against the oracle's
    own 32-step integration (~40 s of CPU)."""
    from covomix_b200.flow import B200FlowSampler
    sd, _ = vosingle
    smp = B200FlowSampler(sd, syn.VOSINGLE, dev, torchdiffeq_ode_method="euler", ode_step_size=1 / 32)
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 1, 650, prompt=150, seed=30)
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
    with torch.inference_mode():
        ref = orc.flow_sample(sd, syn.VOSINGLE, ids, cond, y0, cond_scale=0.7, method="euler", step_size=1 / 32)
    r, ma = rel_l2(mel, ref), float((mel.cpu() - ref).abs().mean())
    record("sample C2 (VoSingle, 32 Euler steps, N=650) vs oracle", rel_l2=r, mean_abs=ma)
    assert r < 3e-2 and ma < 0.05
    smp.close()


def test_vomix_64_euler_steps_vs_oracle(dev, orc, vomix):
    """Error growth over C3's 64 NFE: VoMix, 64 Euler steps, N = 400 (4 key tiles), B = 1, against the oracle."""
    from covomix_b200.flow import B200FlowSampler
    sd, _ = vomix
    smp = B200FlowSampler(sd, syn.VOMIX, dev, torchdiffeq_ode_method="euler", ode_step_size=1 / 64)
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOMIX, 1, 400, prompt=100, seed=33)
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
    with torch.inference_mode():
        ref = orc.flow_sample(sd, syn.VOMIX, ids, cond, y0, cond_scale=0.7, method="euler", step_size=1 / 64)
    r, ma = rel_l2(mel, ref), float((mel.cpu() - ref).abs().mean())
    record("sample VoMix, 64 Euler steps, N=400 vs oracle", rel_l2=r, mean_abs=ma)
    assert r < 3e-2 and ma < 0.05
    smp.close()


def test_step_size_not_dividing_one(dev, orc, vosingle):
    """torchdiffeq's grid for h = 0.3 is 0, .3, .6, .9, 1 (last step shorter), not k/4."""
    from covomix_b200.flow import B200FlowSampler
    sd, _ = vosingle
    smp = B200FlowSampler(sd, syn.VOSINGLE, dev, torchdiffeq_ode_method="midpoint", ode_step_size=0.3)
    assert smp.n_steps() == 4
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 1, 40, prompt=8, seed=5)
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
    with torch.inference_mode():
        ref = orc.flow_sample(sd, syn.VOSINGLE, ids, cond, y0, cond_scale=0.7, method="midpoint", step_size=0.3)
        uniform = orc.flow_sample(sd, syn.VOSINGLE, ids, cond, y0, cond_scale=0.7, method="midpoint", step_size=0.25)
    assert rel_l2(mel, ref) < 1e-2 < rel_l2(uniform, ref)     # the uniform grid is measurably a different answer
    smp.close()


def test_reused_workspace_with_garbage_padding(dev, vosingle, vocoder):
    """a cached plan must not rely on padding columns zeroed at plan-build time -- fill the (library-cached)
    workspaces with NaN bit patterns between two calls of the same shape."""
    _, smp = vosingle
    _, gen = vocoder
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 1, 72, prompt=8, seed=6)
    a = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
    va = smp.velocity(y0.to(dev), times=0.5, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(3), 2, 80, 40).to(dev)
    wa = gen(mel)
    torch.cuda.synchronize()
    with torch.inference_mode():                    # the cached workspaces were allocated under inference_mode
        for ws in list(smp._ws.values()) + list(gen._ws.values()):
            ws.fill_(0xFF)
    b = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
    vb = smp.velocity(y0.to(dev), times=0.5, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    wb = gen(mel)
    assert torch.equal(a, b) and torch.equal(va, vb) and torch.equal(wa, wb)


def test_out_of_range_ids_raise(dev, vosingle):
    _, smp = vosingle
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 1, 16, prompt=4)
    bad = ids.clone()
    bad[0, 3] = 503                     # the table has 503 rows (502 = null id): nn.Embedding raises IndexError
    with pytest.raises(IndexError):
        smp.sample(phoneme_ids=bad.to(dev), cond=cond.to(dev), y0=y0.to(dev))
    bad[0, 3] = -1
    with pytest.raises(IndexError):
        smp.sample(phoneme_ids=bad.to(dev), cond=cond.to(dev), y0=y0.to(dev))


@pytest.mark.parametrize("N", [1, 31, 130])
def test_euler_sample_vs_oracle_ragged_lengths(dev, orc, vosingle, N):
    from covomix_b200.flow import B200FlowSampler
    sd, _ = vosingle
    smp = B200FlowSampler(sd, syn.VOSINGLE, dev, torchdiffeq_ode_method="euler", ode_step_size=0.25)
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 2, N, prompt=min(8, N), seed=N)
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
    ref = orc.flow_sample(sd, syn.VOSINGLE, ids, cond, y0, cond_scale=0.7, method="euler", step_size=0.25)
    assert rel_l2(mel, ref) < 3e-2
    smp.close()


def test_flow_argument_errors(dev, vosingle):
    _, smp = vosingle
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 1, 16, prompt=4)
    with pytest.raises(ValueError):
        smp.sample(phoneme_ids=ids[:, :8].to(dev), cond=cond.to(dev))
    with pytest.raises(ValueError):
        smp.sample(phoneme_ids=ids.to(dev), cond=cond[:, :, :40].to(dev))
    ws = torch.empty(1024, dtype=torch.uint8, device=dev)
    out = torch.empty(1, 16, 80, device=dev)
    rc = nat.lib().covo_flow_sample(smp._h, P(ids.to(dev)), P(cond.to(dev)), P(y0.to(dev)), P(out), 1, 16, 1, 16, 0.7,
                                    P(ws), ws.numel(), None)
    assert rc == -1 and b"workspace too small" in nat.lib().covo_last_error()


# ------------------------------------------------------------------------------------------ flow, full-size properties
def test_c3_batch_items_are_independent(dev, vomix):
    """BASELINE config C3 shape (B=8, N=1650): each item's velocity must equal the B=1 evaluation of that item."""
    _, smp = vomix
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOMIX, 8, 1650, prompt=150, seed=30)
    v = smp.velocity(y0.to(dev), times=0.5, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    assert torch.isfinite(v).all()
    for b in (0, 7):
        vb = smp.velocity(y0[b:b + 1].to(dev), times=0.5, phoneme_ids=ids[b:b + 1].to(dev), cond=cond[b:b + 1].to(dev),
                          cond_scale=0.7)
        assert rel_l2(v[b:b + 1], vb) < 1e-5


def test_cfg_linearity(dev, vosingle):
    """v(s) = (1+s) v_c - s v_n is affine in s: v(0.7) == v(0) + 0.7/0.3 * (v(0.3) - v(0))."""
    _, smp = vosingle
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 1, 200, prompt=50, seed=4)
    f = lambda s: smp.velocity(y0.to(dev), times=0.25, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=s)
    v0, v3, v7 = f(0.0), f(0.3), f(0.7)
    assert rel_l2(v7, v0 + (0.7 / 0.3) * (v3 - v0)) < 1e-5


# ------------------------------------------------------------------------------------------ vocoder
def test_hifigan_matches_reference(dev, vocoder):
    sd, gen = vocoder
    g = np.load(os.path.join(GOLDEN, "hifigan.npz"))
    rng = torch.Generator().manual_seed(int(g["input_seed"]))
    mels = [syn.synthetic_logmel(rng, 1, 80, 256), syn.synthetic_logmel(rng, 80, 64), syn.synthetic_logmel(rng, 2, 80, 48)]
    for mel, key in zip(mels, ("wav_c1", "wav_unbatched", "wav_batch")):
        wav = gen(mel.to(dev))
        ref = torch.from_numpy(g[key])
        assert wav.shape[-1] == 160 * mel.shape[-1] + 32
        assert wav.shape == ref.shape               # reference shapes: [B, 1, L], and [1, L] for the unbatched [80, T] input
        assert rel_l2(wav.reshape(-1), ref.reshape(-1)) < 2e-3
        # mel_decode_to_wav semantics (x32768 -> int16): fused i16 output == host-side cast of our own f32 output
        i16 = gen(mel.to(dev), out_dtype="i16").cpu().numpy().reshape(-1)
        host = (wav.reshape(-1) * 32768.0).cpu().numpy().astype("int16")
        assert np.array_equal(i16, host)
        ri = (ref.reshape(-1) * 32768.0).numpy().astype("int16").astype(np.int64)
        # 16-bit operands give additive noise ~8e-4 of the signal RMS (about -62 dB), not a per-sample relative error
        rms = float(np.sqrt(np.mean(ri.astype(np.float64) ** 2)))
        close = np.abs(i16.astype(np.int64) - ri) <= max(2.0, 4e-3 * rms)
        assert close.mean() > 0.99
        f16 = gen(mel.to(dev), out_dtype="f16")
        assert torch.equal(f16, wav.half())


def test_hifigan_bf16_operands(dev):
    from covomix_b200.vocoder import B200Generator
    g = np.load(os.path.join(GOLDEN, "hifigan.npz"))
    sd = syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234)
    gen = B200Generator(sd, syn.HIFIGAN_COVOMIX, dev, h_format="bf16")
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(int(g["input_seed"])), 1, 80, 256)
    assert rel_l2(gen(mel.to(dev)).reshape(-1), g["wav_c1"].reshape(-1)) < 1e-2
    gen.close()


@pytest.mark.parametrize("fused", ["1", "0"])
def test_hifigan_aligned_polyphase_matches_scatter_path(dev, vocoder, fused):
    """ConvTranspose1d layers with k - 2 pad == stride run in the aligned polyphase form (weights re-indexed at handle creation,
    outputs stored as TMA boxes); COVO_HIFIGAN_SCATTER=1 keeps the classic polyphase form with per-element scatter stores.
    Same products in a different tap order: waveforms agree to fp32 round-off, with and without the fused last stage, for
    batched, ragged and one-frame inputs."""
    from covomix_b200.vocoder import B200Generator
    sd, _ = vocoder
    env = {"COVO_HIFIGAN_NO_FUSED": "0" if fused == "1" else "1"}
    old = {k: os.environ.get(k) for k in ("COVO_HIFIGAN_SCATTER", "COVO_HIFIGAN_NO_FUSED")}
    try:
        os.environ.update(env)
        os.environ["COVO_HIFIGAN_SCATTER"] = "1"
        ref = B200Generator(sd, syn.HIFIGAN_COVOMIX, dev)
        os.environ["COVO_HIFIGAN_SCATTER"] = "0"
        gen = B200Generator(sd, syn.HIFIGAN_COVOMIX, dev)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for B, T in ((3, 300), (1, 129), (2, 1)):
        mel = syn.synthetic_logmel(torch.Generator().manual_seed(40 + T), B, 80, T).to(dev)
        a, b = ref(mel), gen(mel)
        assert a.shape == b.shape and torch.isfinite(b).all()
        assert rel_l2(b, a) < 2e-5, (B, T, rel_l2(b, a))
    ref.close()
    gen.close()


@pytest.mark.parametrize("T", [1, 2, 7])
def test_hifigan_tiny_lengths(dev, orc, vocoder, T):
    sd, gen = vocoder
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(T), 3, 80, T)
    assert rel_l2(gen(mel.to(dev)), orc.hifigan_forward(sd, syn.HIFIGAN_COVOMIX, mel)) < 2e-3


def test_hifigan_resblock2_config(dev, orc):
    """config_v3-style generator: ResBlock2 (models.py:51-72), different rates / kernels / channel counts."""
    from covomix_b200.vocoder import B200Generator
    cfg = syn.HifiganConfig(resblock="2", upsample_rates=(8, 8, 4), upsample_kernel_sizes=(16, 16, 8),
                            upsample_initial_channel=256, resblock_kernel_sizes=(3, 5, 7),
                            resblock_dilation_sizes=((1, 2), (2, 6), (3, 12)))
    sd = syn.synthetic_hifigan_state_dict(cfg, 5)
    gen = B200Generator(sd, cfg, dev)
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(9), 2, 80, 33)
    wav = gen(mel.to(dev))
    ref = orc.hifigan_forward(sd, cfg, mel)
    assert wav.shape == ref.shape == (2, 1, cfg.out_len(33))
    assert rel_l2(wav, ref) < 2e-3
    gen.close()


def test_hifigan_full_size_properties(dev, vocoder):
    """C3's vocoder shape [8, 80, 1500] (and C5's T=4096): items are independent, and the generator is local:
    samples further than the receptive field (~20.4 frames) from a cut are unchanged by the cut."""
    _, gen = vocoder
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(11), 8, 80, 1500).to(dev)
    wav = gen(mel)
    assert wav.shape == (8, 1, 240032) and torch.isfinite(wav).all() and float(wav.abs().max()) <= 1.0
    w3 = gen(mel[3:4])
    assert rel_l2(wav[3:4], w3) < 1e-6
    crop = gen(mel[3:4, :, :700])
    keep = (700 - 24) * 160
    assert rel_l2(crop[..., :keep], w3[..., :keep]) < 1e-6
    long_mel = syn.synthetic_logmel(torch.Generator().manual_seed(12), 1, 80, 4096).to(dev)
    long = gen(long_mel)
    assert long.shape == (1, 1, 160 * 4096 + 32) and torch.isfinite(long).all()
    # time-chunked execution with a receptive-field halo (used for inputs whose workspace would not fit) is exact
    old = gen.MAX_FRAMES_PER_CALL
    try:
        gen.MAX_FRAMES_PER_CALL = 1000
        chunked = gen(long_mel)
    finally:
        gen.MAX_FRAMES_PER_CALL = old
    assert chunked.shape == long.shape and rel_l2(chunked, long) < 1e-6


def test_pipeline_batches_match_single_item_runs(dev, vosingle, vocoder):
    """covomix_b200.pipeline.synthesize (equal-length batching + fused int16) == one utterance at a time."""
    from covomix_b200 import pipeline
    from covomix_b200.flow import B200FlowSampler
    sd, _ = vosingle
    _, gen = vocoder
    smp = B200FlowSampler(sd, syn.VOSINGLE, dev, torchdiffeq_ode_method="euler", ode_step_size=0.25)
    items = []
    for k, (n, p) in enumerate([(96, 16), (64, 8), (96, 20)]):
        ids, cond, y0, mask = syn.synthetic_flow_inputs(syn.VOSINGLE, 1, n, prompt=p, seed=100 + k)
        items.append({"phoneme_ids": ids[0], "cond": cond[0], "mask": mask[0], "y0": y0[0]})
    batched = pipeline.synthesize(smp, gen, items, batch=4)
    for it, w in zip(items, batched):
        single = pipeline.synthesize(smp, gen, [it], batch=1)[0]
        n_gen = int(it["mask"].sum())
        assert w.dtype == np.int16 and w.shape == (160 * n_gen + 32,) and single.shape == w.shape
        d = np.abs(w.astype(np.int64) - single.astype(np.int64))
        assert d.max() <= 2                      # batch composition only changes GEMM tile scheduling, not arithmetic order
    smp.close()


# ------------------------------------------------------------------------------------------ optional execution paths
def _fresh_sampler(dev, sd, cfg, env, **kw):
    """A sampler created under extra environment settings (the library reads them at handle creation)."""
    from covomix_b200.flow import B200FlowSampler
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return B200FlowSampler(sd, cfg, dev, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("method,step", [("euler", 1 / 8), ("midpoint", 0.25)])
def test_persistent_step_kernel_matches_launch_per_op_path(dev, vosingle, method, step):
    """COVO_FLOW_PERSISTENT=1: the whole ODE loop in one cooperative launch (csrc/flow_persistent.cuh) walks the same
    tile loops as the launch-per-op graph, so the sampled mel must agree to fp32 round-off of the tile schedule."""
    sd, _ = vosingle
    ref_smp = _fresh_sampler(dev, sd, syn.VOSINGLE, {"COVO_FLOW_PERSISTENT": "0"}, torchdiffeq_ode_method=method, ode_step_size=step)
    per_smp = _fresh_sampler(dev, sd, syn.VOSINGLE, {"COVO_FLOW_PERSISTENT": "1"}, torchdiffeq_ode_method=method, ode_step_size=step)
    for B, N in ((1, 200), (2, 129)):
        ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, B, N, prompt=20, seed=N)
        a = ref_smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
        b = per_smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
        assert per_smp.last_launches() == 4 and ref_smp.last_launches() > 100
        assert torch.isfinite(b).all() and rel_l2(b, a) < 1e-5
    ref_smp.close()
    per_smp.close()


def test_serpentine_row_order_is_pure_scheduling(dev, vomix):
    """COVO_FLOW_SERPENTINE (default on): every kernel of the velocity net walks the token rows in the opposite direction of
    its producer so that it starts on the rows still in L2.  Only the order of tiles / items / rows changes, never the
    arithmetic inside one: results are bit-identical to the ascending order."""
    sd, _ = vomix
    up = _fresh_sampler(dev, sd, syn.VOMIX, {"COVO_FLOW_SERPENTINE": "0"}, torchdiffeq_ode_method="euler", ode_step_size=0.25)
    sp = _fresh_sampler(dev, sd, syn.VOMIX, {"COVO_FLOW_SERPENTINE": "1"}, torchdiffeq_ode_method="euler", ode_step_size=0.25)
    for B, N in ((3, 300), (1, 129), (2, 1)):
        ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOMIX, B, N, prompt=min(30, N), seed=12)
        a = up.velocity(y0.to(dev), times=0.3, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
        b = sp.velocity(y0.to(dev), times=0.3, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
        assert torch.equal(a, b)
        sa = up.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
        sb = sp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7, y0=y0.to(dev))
        assert torch.isfinite(sb).all() and torch.equal(sa, sb)
    up.close()
    sp.close()


def test_cluster_pair_gemm_path_matches(dev, vosingle):
    """COVO_GEMM_MC=2: GEMMs as cluster pairs with a multicast weight tile (gemm_tc_pair_kernel) -- same arithmetic."""
    sd, smp = vosingle
    mc_smp = _fresh_sampler(dev, sd, syn.VOSINGLE, {"COVO_GEMM_MC": "2"})
    ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, 2, 300, prompt=30, seed=8)
    a = smp.velocity(y0.to(dev), times=0.3, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    b = mc_smp.velocity(y0.to(dev), times=0.3, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    assert rel_l2(b, a) < 1e-5
    mc_smp.close()


def test_cta_pair_mma_gemm_path_matches(dev, vosingle):
    """COVO_GEMM_CG=2 forces every GEMM with >= 2 M tiles onto the cta_group::2 kernel (gemm_tc_cg2_kernel: one M = 256 MMA per
    CTA pair, each CTA staging half of the weight tile; the default uses it only for multi-wave GEMMs), COVO_GEMM_CG=1 forbids
    it.  Same products, same K order -> same results.  N = 330 gives 11 M tiles: the last pair has a phantom second tile."""
    sd, _ = vosingle
    one = _fresh_sampler(dev, sd, syn.VOSINGLE, {"COVO_GEMM_CG": "1"})
    two = _fresh_sampler(dev, sd, syn.VOSINGLE, {"COVO_GEMM_CG": "2"})
    for B, N in ((2, 330), (1, 129)):
        ids, cond, y0, _ = syn.synthetic_flow_inputs(syn.VOSINGLE, B, N, prompt=30, seed=9)
        a = one.velocity(y0.to(dev), times=0.3, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
        b = two.velocity(y0.to(dev), times=0.3, phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
        assert torch.isfinite(b).all() and rel_l2(b, a) < 1e-5
    one.close()
    two.close()
