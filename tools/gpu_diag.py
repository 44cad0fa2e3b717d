"""On-GPU diagnostics: each check prints error norms instead of only asserting.  Run one check per
process (a CUDA fault poisons the context):  python tools/gpu_diag.py <check> [args]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import covomix_b200  # noqa
from covomix_b200 import _native as nat, synthetic as syn  # noqa
from covomix_b200.flow import B200FlowSampler  # noqa
from covomix_b200.vocoder import B200Generator  # noqa
from oracle import covomix_oracle as orc  # noqa

dev = torch.device("cuda:0")
P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)), float((a - b).abs().max())


def check_gemm():
    L = nat.lib()
    torch.manual_seed(0)
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 256, 64, 256), (256, 256, 128, 128), (300, 512, 1024, 0),
                          (1300, 1024, 1024, 0), (4096, 3072, 1024, 256), (5000, 1024, 4096, 256)]:
        A = torch.randn(M, K, device=dev).bfloat16()
        W = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        out = torch.full((M, N), float("nan"), device=dev)
        outh = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        nat.check(L.covo_dbg_gemm(P(A), P(W), P(bias), P(res), P(out), P(outh), M, N, K, 1, bn, None), "dbg_gemm")
        torch.cuda.synchronize()
        ref = A.float() @ W.float().t() + bias + res
        r = rel(out, ref)
        rh = rel(outh.float(), torch.nn.functional.gelu(ref))
        print(f"gemm M={M} N={N} K={K} bn={bn}: rel={r[0]:.3e} maxabs={r[1]:.3e} | gelu-bf16 rel={rh[0]:.3e} nan={int(out.isnan().sum())}", flush=True)


def check_attn():
    L = nat.lib()
    torch.manual_seed(0)
    for (Bt, N, H) in [(1, 128, 1), (1, 256, 2), (1, 200, 1), (2, 650, 16), (2, 1650, 16)]:
        qkv = torch.randn(Bt, N, 3 * H * 64, device=dev).bfloat16()
        o1 = torch.zeros(Bt, N, H * 64, device=dev, dtype=torch.bfloat16)
        nat.check(L.covo_dbg_attention(P(qkv), P(o1), Bt, N, H, 0, None), "dbg_attention")
        torch.cuda.synchronize()
        q, k, v = (t.reshape(Bt, N, H, 64).permute(0, 2, 1, 3).float() for t in qkv.chunk(3, dim=-1))
        ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
        ref = ref.permute(0, 2, 1, 3).reshape(Bt, N, H * 64)
        r = rel(o1.float(), ref)
        print(f"attn Bt={Bt} N={N} H={H}: tc rel={r[0]:.3e} maxabs={r[1]:.3e} nan={int(o1.float().isnan().sum())}", flush=True)
        if N <= 650:
            o0 = torch.zeros_like(o1)
            nat.check(L.covo_dbg_attention(P(qkv), P(o0), Bt, N, H, 1, None), "dbg_attention naive")
            torch.cuda.synchronize()
            r = rel(o0.float(), ref)
            print(f"      naive rel={r[0]:.3e}", flush=True)


def check_hifigan(fmt="bf16"):
    cfg = syn.HIFIGAN_COVOMIX
    sd = syn.synthetic_hifigan_state_dict(cfg, 1234)
    gen = B200Generator(sd, cfg, dev, h_format=fmt)
    g = torch.Generator().manual_seed(30)
    for shape in [(1, 80, 256), (2, 80, 48), (80, 64)]:
        mel = syn.synthetic_logmel(g, *shape)
        ref = orc.hifigan_forward(sd, cfg, mel)
        t0 = time.time()
        wav = gen(mel.to(dev))
        torch.cuda.synchronize()
        r = rel(wav, ref)
        i16 = gen(mel.to(dev), out_dtype="i16").cpu().numpy().reshape(-1).astype(np.int64)
        ri = orc.wav_to_int16(ref).reshape(-1).astype(np.int64)
        close = np.mean(np.abs(i16 - ri) <= np.maximum(2, np.abs(ri) / 256))
        print(f"hifigan[{fmt}] {shape}: rel={r[0]:.3e} maxabs={r[1]:.3e} shape={tuple(wav.shape)} int16-close={close:.4f} ({time.time()-t0:.3f}s)", flush=True)


def check_flow(name="vosingle", naive="0"):
    os.environ["COVO_DEBUG_NAIVE_ATTN"] = naive
    cfg = syn.VOSINGLE if name == "vosingle" else syn.VOMIX
    g = np.load(os.path.join(ROOT, "tests", "golden", f"flow_{name}.npz"))
    sd = syn.synthetic_flow_state_dict(cfg, int(g["weight_seed"]))
    B, N = int(g["B"]), int(g["N"])
    ids, cond, y0, mask = syn.synthetic_flow_inputs(cfg, B, N, prompt=int(g["prompt"]), seed=int(g["input_seed"]))
    smp = B200FlowSampler(sd, cfg, dev)
    v = smp.velocity(y0.to(dev), times=float(g["t"]), phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=0.7)
    torch.cuda.synchronize()
    r = rel(v, torch.from_numpy(g["v_cfg"]))
    print(f"flow[{name}] naive_attn={naive} velocity vs reference: rel={r[0]:.3e} maxabs={r[1]:.3e} std_ref={float(g['v_cfg'].std()):.3f} nan={int(v.isnan().sum())}", flush=True)
    v1 = smp.velocity(y0.to(dev), times=float(g["t"]), phoneme_ids=ids.to(dev), cond=cond.to(dev), cond_scale=1.0)
    r = rel(v1, torch.from_numpy(g["v_cond"]))
    print(f"   cond branch only: rel={r[0]:.3e}", flush=True)
    t0 = time.time()
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), mask=mask.to(dev), cond_scale=0.7,
                     y0=torch.from_numpy(g["y0_sample"]).to(dev))
    torch.cuda.synchronize()
    t1 = time.time()
    r = rel(mel, torch.from_numpy(g["mel"]))
    ma = float((mel.cpu() - torch.from_numpy(g["mel"])).abs().mean())
    print(f"   sample (midpoint/16) vs reference: rel={r[0]:.3e} maxabs={r[1]:.3e} meanabs={ma:.3e} ({t1-t0:.2f}s incl. plan build)", flush=True)
    t0 = time.time()
    mel = smp.sample(phoneme_ids=ids.to(dev), cond=cond.to(dev), mask=mask.to(dev), cond_scale=0.7,
                     y0=torch.from_numpy(g["y0_sample"]).to(dev))
    torch.cuda.synchronize()
    print(f"   second sample call: {time.time()-t0:.3f}s", flush=True)


if __name__ == "__main__":
    fn = globals()["check_" + sys.argv[1]]
    fn(*sys.argv[2:])
