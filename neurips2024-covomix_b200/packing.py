"""Checkpoint -> packed weight blob for libcovomix_b200.so.

Input is a state dict in the reference's own key layout:
``CoVoMix.state_dict()`` keys, optionally prefixed by ``CoVoMix.`` / ``cfm_wrapper.CoVoMix.``
(Lightning checkpoint, covomix/conditional_model.py:113), and ``Generator.state_dict()`` keys
with or without weight norm (``weight_g`` / ``weight_v``; hifi-gan/models.py:118-125).

Blob layout (little endian; parsed by csrc/common.cuh ``Weights::load``):
    header : magic "COVOWTS1", u32 n_entries, u32 reserved
    entries: n x { char name[48]; u32 dtype; u32 ndim; u64 shape[4]; u64 offset; u64 nbytes }
    data   : 256-byte aligned tensors
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np
import torch

from .synthetic import FlowConfig, HifiganConfig, T2SConfig

DT_F32, DT_BF16, DT_F16, DT_I64 = 0, 1, 2, 3
_ENTRY = struct.Struct("<48sII4QQQ")


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class BlobBuilder:
    def __init__(self):
        self.items: List[Tuple[str, int, Tuple[int, ...], bytes]] = []

    def add(self, name: str, t: torch.Tensor, dtype: int):
        assert len(name) < 48, name
        t = t.detach().cpu().contiguous()
        if dtype == DT_F32:
            raw = t.to(torch.float32).numpy().tobytes()
        elif dtype == DT_BF16:
            raw = t.to(torch.bfloat16).view(torch.int16).numpy().tobytes()
        elif dtype == DT_F16:
            raw = t.to(torch.float16).view(torch.int16).numpy().tobytes()
        else:
            raise ValueError(dtype)
        assert t.ndim <= 4
        self.items.append((name, dtype, tuple(t.shape), raw))

    def build(self) -> np.ndarray:
        n = len(self.items)
        off = _round_up(16 + n * _ENTRY.size, 256)
        entries, chunks = [], []
        for name, dtype, shape, raw in self.items:
            shp = list(shape) + [0] * (4 - len(shape))
            entries.append(_ENTRY.pack(name.encode(), dtype, len(shape), *shp, off, len(raw)))
            chunks.append((off, raw))
            off = _round_up(off + len(raw), 256)
        blob = np.zeros(off, dtype=np.uint8)
        blob[:8] = np.frombuffer(b"COVOWTS1", dtype=np.uint8)
        blob[8:16] = np.frombuffer(struct.pack("<II", n, 0), dtype=np.uint8)
        p = 16
        for e in entries:
            blob[p:p + len(e)] = np.frombuffer(e, dtype=np.uint8)
            p += len(e)
        for o, raw in chunks:
            blob[o:o + len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        return blob


def blob_entry_names(blob: np.ndarray) -> List[str]:
    """Names of the tensors in a packed blob (the table the C side binds by name, csrc/common.cuh ``Weights``)."""
    n = struct.unpack("<I", bytes(blob[8:12]))[0]
    out = []
    for i in range(n):
        rec = _ENTRY.unpack(bytes(blob[16 + i * _ENTRY.size:16 + (i + 1) * _ENTRY.size]))
        out.append(rec[0].split(b"\0", 1)[0].decode())
    return out


# --------------------------------------------------------------------------------------
# flow-matching velocity net
# --------------------------------------------------------------------------------------

def strip_flow_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accept Lightning (``cfm_wrapper.CoVoMix.*``), wrapper (``CoVoMix.*``) or bare keys."""
    for prefix in ("cfm_wrapper.CoVoMix.", "CoVoMix."):
        if any(k.startswith(prefix) for k in sd):
            return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return dict(sd)


def flow_config_from_state_dict(sd: Dict[str, torch.Tensor], heads: int = 16, dim_head: int = 64) -> FlowConfig:
    """Recover the ``CoVoMix(...)`` constructor arguments from tensor shapes."""
    sd = strip_flow_prefix(sd)
    dim = sd["to_embed.weight"].shape[0]
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    dim_in = sd["null_cond"].shape[0]
    n_tok, demb = sd["to_phoneme_emb.weight"].shape
    dim_x = sd["to_pred.weight"].shape[0]
    embed_in = sd["to_embed.weight"].shape[1]
    two = embed_in == dim_in + 80 + 2 * demb and dim_x == 80 and dim_in != 80
    return FlowConfig(dim=dim, depth=depth, heads=heads, dim_head=dim_head, dim_in=dim_in, num_phoneme_tokens=n_tok - 1,
                      dim_phoneme_emb=demb, conv_pos_kernel=sd["conv_embed.dw_conv1d.0.weight"].shape[-1],
                      twocondition_oneoutput=two)


def pack_flow_weights(sd: Dict[str, torch.Tensor], cfg: FlowConfig) -> np.ndarray:
    sd = {k: v.detach().float().cpu() for k, v in strip_flow_prefix(sd).items()}
    D, dx, S, demb = cfg.dim, cfg.dim_x, cfg.n_streams, cfg.dim_phoneme_emb
    b = BlobBuilder()
    b.add("null_cond", sd["null_cond"], DT_F32)
    b.add("time.w", sd["sinu_pos_emb.0.weights"], DT_F32)
    b.add("time.lin.w", sd["sinu_pos_emb.1.weight"], DT_F32)
    b.add("time.lin.b", sd["sinu_pos_emb.1.bias"], DT_F32)
    b.add("emb.table", sd["to_phoneme_emb.weight"], DT_F32)
    # to_embed acts on cat(x, phoneme_emb, cond) (acoustic.py:503-505): split its columns into the
    # per-evaluation part (x) and the per-call constant part (emb | cond).
    W = sd["to_embed.weight"]
    assert W.shape[1] == cfg.embed_in == dx + S * demb + cfg.dim_in, (W.shape, cfg)
    ldx = _round_up(dx, 64)
    kpc = _round_up(S * demb + cfg.dim_in, 64)
    wx = torch.zeros(D, ldx)
    wx[:, :dx] = W[:, :dx]
    wpc = torch.zeros(D, kpc)
    wpc[:, :S * demb + cfg.dim_in] = W[:, dx:]
    b.add("embed.wx", wx, DT_BF16)
    b.add("embed.wpc", wpc, DT_BF16)
    b.add("embed.b", sd["to_embed.bias"], DT_F32)
    b.add("convpos.wT", sd["conv_embed.dw_conv1d.0.weight"][:, 0, :].t().contiguous(), DT_F32)   # [31, D]
    b.add("convpos.b", sd["conv_embed.dw_conv1d.0.bias"], DT_F32)
    b.add("rope.inv_freq", sd["transformer.rotary_emb.inv_freq"], DT_F32)
    gw, gb = [], []
    for L in range(cfg.depth):
        for norm in ("1", "3"):
            for part in ("to_gamma", "to_beta"):
                gw.append(sd[f"transformer.layers.{L}.{norm}.{part}.weight"])
                gb.append(sd[f"transformer.layers.{L}.{norm}.{part}.bias"])
    b.add("adaln.w", torch.cat(gw, 0), DT_F32)               # [depth*4*D, 4D]
    b.add("adaln.b", torch.cat(gb, 0), DT_F32)
    b.add("final.gamma", sd["transformer.final_norm.gamma"], DT_F32)
    npred = _round_up(dx, 64)
    pw = torch.zeros(npred, D)
    pw[:dx] = sd["to_pred.weight"]
    b.add("pred.w", pw, DT_BF16)
    for L in range(cfg.depth):
        p = f"transformer.layers.{L}."
        if L >= cfg.depth // 2:
            b.add(f"L{L}.skip.w", sd[p + "0.weight"], DT_BF16)
            b.add(f"L{L}.skip.b", sd[p + "0.bias"], DT_F32)
        b.add(f"L{L}.qkv.w", sd[p + "2.to_qkv.weight"], DT_BF16)
        b.add(f"L{L}.out.w", sd[p + "2.to_out.weight"], DT_BF16)
        b.add(f"L{L}.ff1.w", sd[p + "4.0.weight"], DT_BF16)
        b.add(f"L{L}.ff1.b", sd[p + "4.0.bias"], DT_F32)
        b.add(f"L{L}.ff2.w", sd[p + "4.2.weight"], DT_BF16)
        b.add(f"L{L}.ff2.b", sd[p + "4.2.bias"], DT_F32)
    return b.build()


# --------------------------------------------------------------------------------------
# HiFi-GAN generator
# --------------------------------------------------------------------------------------

def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``remove_weight_norm`` (hifi-gan/models.py:118-125): w = g * v / ||v|| with the norm over all
    dims but 0 (torch.nn.utils.weight_norm default dim=0, also for ConvTranspose1d)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[:-len(".weight_g")]
            g, vv = v.float(), sd[base + ".weight_v"].float()
            norm = vv.flatten(1).norm(dim=1).reshape(-1, *([1] * (vv.ndim - 1)))
            out[base + ".weight"] = vv * (g / norm)
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def pack_hifigan_weights(sd: Dict[str, torch.Tensor], cfg: HifiganConfig, h_format: str = "bf16") -> np.ndarray:
    if "generator" in sd and isinstance(sd["generator"], dict):        # vocoder ckpt: {'generator': state_dict}
        sd = sd["generator"]
    sd = {k: v.detach().float().cpu() for k, v in fold_weight_norm(sd).items()}
    hdt = DT_F16 if h_format == "fp16" else DT_BF16
    pad = lambda c: _round_up(c, 64)
    b = BlobBuilder()
    c0 = cfg.upsample_initial_channel

    def pack_conv(w: torch.Tensor, cin_pad: int, cout_pad: int) -> torch.Tensor:
        co, ci, k = w.shape                                           # Conv1d weight [Cout, Cin, K]
        out = torch.zeros(cout_pad, k, cin_pad)
        out[:co, :, :ci] = w.permute(0, 2, 1)
        return out.reshape(cout_pad, k * cin_pad)

    def pad_vec(v: torch.Tensor, n: int) -> torch.Tensor:
        out = torch.zeros(n)
        out[:v.numel()] = v
        return out

    b.add("conv_pre.w", pack_conv(sd["conv_pre.weight"], pad(cfg.num_mels), pad(c0)), hdt)
    b.add("conv_pre.b", pad_vec(sd["conv_pre.bias"], pad(c0)), DT_F32)
    nk = len(cfg.resblock_kernel_sizes)
    ch = c0
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, cout = c0 // 2 ** i, c0 // 2 ** (i + 1)
        cip, cop = pad(cin), pad(cout)
        w = sd[f"ups.{i}.weight"]                                     # ConvTranspose1d weight [Cin, Cout, K]
        assert tuple(w.shape) == (cin, cout, k), (w.shape, cin, cout, k)
        J = (k + u - 1) // u
        # polyphase: row r*Cop + co, column j*Cip + ci  <-  W[ci, co, r + u*j]
        pw = torch.zeros(u, cop, J, cip)
        for r in range(u):
            for j in range(J):
                kk = r + u * j
                if kk < k:
                    pw[r, :cout, j, :cin] = w[:, :, kk].t()
        b.add(f"ups.{i}.w", pw.reshape(u * cop, J * cip), hdt)
        b.add(f"ups.{i}.b", pad_vec(sd[f"ups.{i}.bias"], cop).repeat(u), DT_F32)
        ch = cout
        for j in range(nk):
            r = i * nk + j
            names = (("convs1", "c1"), ("convs2", "c2")) if cfg.resblock == "1" else (("convs", "c"),)
            for src, dst in names:
                for m in range(len(cfg.resblock_dilation_sizes[j])):
                    b.add(f"rb.{r}.{dst}.{m}.w", pack_conv(sd[f"resblocks.{r}.{src}.{m}.weight"], cop, cop), hdt)
                    b.add(f"rb.{r}.{dst}.{m}.b", pad_vec(sd[f"resblocks.{r}.{src}.{m}.bias"], cop), DT_F32)
    wpost = sd["conv_post.weight"]                                    # [1, ch, 7]
    pp = torch.zeros(7, pad(ch))
    pp[:, :ch] = wpost[0].t()
    b.add("conv_post.w", pp, DT_F32)
    b.add("conv_post.b", sd["conv_post.bias"].reshape(1), DT_F32)
    return b.build()


# --------------------------------------------------------------------------------------
# text-to-semantic (CoSingle / CoMix)
# --------------------------------------------------------------------------------------

def strip_t2s_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accept Lightning (``cfm_wrapper.model.*``, conditional_model.py:136), wrapper (``model.*``) or bare keys."""
    for prefix in ("cfm_wrapper.model.", "model."):
        if any(k.startswith(prefix) for k in sd):
            return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return dict(sd)


def t2s_config_from_state_dict(sd: Dict[str, torch.Tensor], heads: int = 8, dim_head: int = 64) -> T2SConfig:
    """Recover the ``TextToSemantic(...)`` arguments (conditional_model.py:122-135) from tensor shapes."""
    sd = strip_t2s_prefix(sd)
    n_text, dim = sd["token_emb.text.weight"].shape
    n_sem, demb = sd["token_emb.speech.weight"].shape
    dt = sd["start_token.speech"].shape[0]
    depth = lambda side: 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(side + "_transformer.layers."))
    return T2SConfig(dim=dim, source_depth=depth("source"), target_depth=depth("target"), heads=heads, dim_head=dim_head,
                     num_text_token_ids=n_text - 1, num_semantic_token_ids=n_sem - 1, two_output=dt == 2 * demb,
                     target_transformer_dim=dt)


def pack_t2s_weights(sd: Dict[str, torch.Tensor], cfg: T2SConfig, weight_format: str = "bf16") -> np.ndarray:
    """Source side in fp32 (runs once per call); the decoder matrices the persistent decode kernel streams every step in
    bf16 (default) or fp32; norms, biases, the tied semantic embedding / logit table and the cross-attention k/v
    projection (applied once per call) in fp32."""
    sd = {k: v.detach().float().cpu() for k, v in strip_t2s_prefix(sd).items()}
    wdt = DT_F32 if weight_format == "fp32" else DT_BF16
    b = BlobBuilder()
    b.add("enc.emb", sd["token_emb.text.weight"], DT_F32)
    b.add("rope.inv_freq", sd["target_transformer.layers.0.0.rotary_emb.freqs"], DT_F32)
    for L in range(cfg.source_depth):
        p, q = f"source_transformer.layers.{L}.", f"enc.L{L}."
        b.add(q + "attn.gamma", sd[p + "0.norm.gamma"], DT_F32)
        b.add(q + "q.w", sd[p + "0.to_q.0.weight"], DT_F32)
        b.add(q + "kv.w", sd[p + "0.to_kv.0.weight"], DT_F32)
        b.add(q + "out.w", sd[p + "0.to_out.weight"], DT_F32)
        b.add(q + "ff.gamma", sd[p + "2.0.gamma"], DT_F32)
        b.add(q + "ff1.w", sd[p + "2.1.weight"], DT_F32)
        b.add(q + "ff1.b", sd[p + "2.1.bias"], DT_F32)
        b.add(q + "ff2.w", sd[p + "2.4.weight"], DT_F32)
        b.add(q + "ff2.b", sd[p + "2.4.bias"], DT_F32)
    b.add("enc.final.gamma", sd["source_transformer.final_norm.gamma"] if cfg.source_depth else torch.ones(cfg.dim), DT_F32)
    b.add("dec.emb", sd["token_emb.speech.weight"], DT_F32)
    b.add("dec.start", sd["start_token.speech"], DT_F32)
    b.add("dec.final.gamma", sd["target_transformer.final_norm.gamma"], DT_F32)
    fi = cfg.ff_inner(cfg.target_transformer_dim)
    fip = _round_up(fi, 64)          # whole 64-k blocks for the tensor-core path of the decode kernel
    for L in range(cfg.target_depth):
        p, q = f"target_transformer.layers.{L}.", f"dec.L{L}."
        b.add(q + "sa.gamma", sd[p + "0.norm.gamma"], DT_F32)
        b.add(q + "sa.qkv.w", torch.cat((sd[p + "0.to_q.0.weight"], sd[p + "0.to_kv.0.weight"]), 0), wdt)   # q | k | v rows
        b.add(q + "sa.out.w", sd[p + "0.to_out.weight"], wdt)
        b.add(q + "ca.gamma", sd[p + "1.norm.gamma"], DT_F32)
        b.add(q + "ca.q.w", sd[p + "1.to_q.0.weight"], wdt)
        b.add(q + "ca.kv.w", sd[p + "1.to_kv.0.weight"], DT_F32)
        b.add(q + "ca.null_kv", sd[p + "1.null_kv"].reshape(2, cfg.heads, cfg.dim_head), DT_F32)
        b.add(q + "ca.out.w", sd[p + "1.to_out.weight"], wdt)
        b.add(q + "ff.gamma", sd[p + "2.0.gamma"], DT_F32)
        b.add(q + "ff1.w", sd[p + "2.1.weight"], wdt)
        b.add(q + "ff1.b", sd[p + "2.1.bias"], DT_F32)
        w2 = torch.zeros(cfg.target_transformer_dim, fip)
        w2[:, :fi] = sd[p + "2.4.weight"]
        b.add(q + "ff2.w", w2, wdt)
        b.add(q + "ff2.b", sd[p + "2.4.bias"], DT_F32)
    return b.build()
