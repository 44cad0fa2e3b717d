"""Seeded synthetic checkpoints and inputs for the CoVoMix hot path.

The reference ships no trained checkpoints (reference README.md:30), so parity and
throughput are measured on random-init weights of the reference architecture.  The
reference's default init is degenerate for that purpose:

* ``AdaptiveRMSNorm`` starts as the identity (covomix/covomix_model/acoustic.py:190-196),
  so time conditioning would be a no-op;
* ``null_cond`` is all zeros (acoustic.py:382);
* HiFi-GAN convs are drawn from N(0, 0.01) (hifi-gan/utils.py:22-25), which gives a
  waveform of std ~5e-3.

The generators below draw every tensor from one ``torch.Generator`` in a fixed key
order, so the same (config, seed) reproduces the same state dict on any machine.  The
key layout is exactly the reference's ``state_dict()`` layout, so
the dicts load into the reference modules with ``load_state_dict`` (that is how
tests/golden/make_golden.py pins the oracle).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


# --------------------------------------------------------------------------------------
# configs (mirror the reference constructors)
# --------------------------------------------------------------------------------------

@dataclass(frozen=True)
class FlowConfig:
    """Arguments of ``CoVoMix(...)`` (acoustic.py:326-348) as used by
    conditional_model.py:99-115 with running_command/Acous_Vo{Single,Mix}.sh."""
    dim: int = 1024
    depth: int = 8
    heads: int = 16
    dim_head: int = 64
    dim_in: int = 80                    # 80 VoSingle, 160 VoMix
    num_phoneme_tokens: int = 502
    dim_phoneme_emb: int = 1024
    ff_mult: int = 4
    conv_pos_kernel: int = 31
    twocondition_oneoutput: bool = False

    @property
    def n_streams(self) -> int:
        return 2 if self.twocondition_oneoutput else 1

    @property
    def dim_x(self) -> int:             # width of the ODE state / of to_pred
        return 80 if self.twocondition_oneoutput else self.dim_in

    @property
    def embed_in(self) -> int:          # acoustic.py:375-380
        if self.twocondition_oneoutput:
            return self.dim_in + 80 + 2 * self.dim_phoneme_emb
        return self.dim_in * 2 + self.dim_phoneme_emb

    @property
    def time_hidden(self) -> int:
        return self.dim * 4


VOSINGLE = FlowConfig(dim_in=80)
VOMIX = FlowConfig(dim_in=160, twocondition_oneoutput=True)


@dataclass(frozen=True)
class HifiganConfig:
    """hifi-gan/config_covomix.json (the fields Generator.__init__ reads, models.py:76-98)."""
    resblock: str = "1"
    upsample_rates: Tuple[int, ...] = (5, 4, 4, 2)
    upsample_kernel_sizes: Tuple[int, ...] = (8, 8, 4, 4)
    upsample_initial_channel: int = 500
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    num_mels: int = 80

    @property
    def hop(self) -> int:
        return int(math.prod(self.upsample_rates))

    def out_len(self, T: int) -> int:
        L = T
        for u, k in zip(self.upsample_rates, self.upsample_kernel_sizes):
            p = (k - u) // 2
            L = (L - 1) * u - 2 * p + k
        return L

    @classmethod
    def from_json(cls, d: dict) -> "HifiganConfig":
        return cls(
            resblock=str(d["resblock"]),
            upsample_rates=tuple(d["upsample_rates"]),
            upsample_kernel_sizes=tuple(d["upsample_kernel_sizes"]),
            upsample_initial_channel=int(d["upsample_initial_channel"]),
            resblock_kernel_sizes=tuple(d["resblock_kernel_sizes"]),
            resblock_dilation_sizes=tuple(tuple(x) for x in d["resblock_dilation_sizes"]),
            num_mels=int(d.get("num_mels", 80)),
        )


HIFIGAN_COVOMIX = HifiganConfig()


# --------------------------------------------------------------------------------------
# flow-matching velocity net: state dict in the reference's key layout
# --------------------------------------------------------------------------------------

def _randn(gen: torch.Generator, *shape: int, std: float = 1.0) -> torch.Tensor:
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std


def synthetic_flow_state_dict(cfg: FlowConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """State dict of ``CoVoMix`` (keys as in ``CoVoMix.state_dict()``), non-degenerate."""
    g = torch.Generator().manual_seed(seed)
    d, th = cfg.dim, cfg.time_hidden
    sd: Dict[str, torch.Tensor] = {}
    sd["null_cond"] = _randn(g, cfg.dim_in)                              # acoustic.py:382 (zeros there)
    sd["sinu_pos_emb.0.weights"] = _randn(g, d // 2)                      # acoustic.py:105
    sd["sinu_pos_emb.1.weight"] = _randn(g, th, d, std=d ** -0.5)
    sd["sinu_pos_emb.1.bias"] = _randn(g, th, std=0.02)
    sd["to_phoneme_emb.weight"] = _randn(g, cfg.num_phoneme_tokens + 1, cfg.dim_phoneme_emb)
    sd["to_embed.weight"] = _randn(g, d, cfg.embed_in, std=cfg.embed_in ** -0.5)
    sd["to_embed.bias"] = _randn(g, d, std=0.02)
    sd["conv_embed.dw_conv1d.0.weight"] = _randn(g, d, 1, cfg.conv_pos_kernel, std=cfg.conv_pos_kernel ** -0.5)
    sd["conv_embed.dw_conv1d.0.bias"] = _randn(g, d, std=0.02)
    sd["transformer.rotary_emb.inv_freq"] = 1.0 / (
        10000 ** (torch.arange(0, cfg.dim_head, 2).float() / cfg.dim_head))   # acoustic.py:119
    inner = cfg.heads * cfg.dim_head
    for L in range(cfg.depth):
        p = f"transformer.layers.{L}."
        if L + 1 > cfg.depth // 2:                                        # acoustic.py:276-279
            sd[p + "0.weight"] = _randn(g, d, 2 * d, std=(2 * d) ** -0.5)
            sd[p + "0.bias"] = _randn(g, d, std=0.02)
        for norm in ("1", "3"):                                           # AdaptiveRMSNorm, acoustic.py:187-196
            sd[p + norm + ".to_gamma.weight"] = _randn(g, d, th, std=0.02)
            sd[p + norm + ".to_gamma.bias"] = 1.0 + _randn(g, d, std=0.1)
            sd[p + norm + ".to_beta.weight"] = _randn(g, d, th, std=0.02)
            sd[p + norm + ".to_beta.bias"] = _randn(g, d, std=0.1)
        sd[p + "2.to_qkv.weight"] = _randn(g, 3 * inner, d, std=d ** -0.5)
        sd[p + "2.to_out.weight"] = _randn(g, d, inner, std=inner ** -0.5)
        sd[p + "4.0.weight"] = _randn(g, d * cfg.ff_mult, d, std=d ** -0.5)
        sd[p + "4.0.bias"] = _randn(g, d * cfg.ff_mult, std=0.02)
        sd[p + "4.2.weight"] = _randn(g, d, d * cfg.ff_mult, std=(d * cfg.ff_mult) ** -0.5)
        sd[p + "4.2.bias"] = _randn(g, d, std=0.02)
    sd["transformer.final_norm.gamma"] = 1.0 + _randn(g, d, std=0.1)
    sd["to_pred.weight"] = _randn(g, cfg.dim_x, d, std=d ** -0.5)
    return sd


# --------------------------------------------------------------------------------------
# HiFi-GAN generator: state dict after remove_weight_norm (plain weight/bias keys)
# --------------------------------------------------------------------------------------

def hifigan_layer_shapes(cfg: HifiganConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key prefix, weight shape) of every conv of ``Generator`` in state_dict order
    (hifi-gan/models.py:76-98)."""
    c0 = cfg.upsample_initial_channel
    out: List[Tuple[str, Tuple[int, ...]]] = [("conv_pre", (c0, cfg.num_mels, 7))]
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        out.append((f"ups.{i}", (c0 // 2 ** i, c0 // 2 ** (i + 1), k)))   # ConvTranspose1d [Cin,Cout,K]
    ch = c0
    for i in range(len(cfg.upsample_rates)):
        ch = c0 // 2 ** (i + 1)
        for j, (k, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            r = i * len(cfg.resblock_kernel_sizes) + j
            if cfg.resblock == "1":
                for m in range(len(dil)):
                    out.append((f"resblocks.{r}.convs1.{m}", (ch, ch, k)))
                for m in range(len(dil)):
                    out.append((f"resblocks.{r}.convs2.{m}", (ch, ch, k)))
            else:
                for m in range(len(dil)):
                    out.append((f"resblocks.{r}.convs.{m}", (ch, ch, k)))
    out.append(("conv_post", (1, ch, 7)))
    return out


def synthetic_hifigan_state_dict(cfg: HifiganConfig = HIFIGAN_COVOMIX, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Generator weights (post ``remove_weight_norm``) scaled so activations stay O(1):
    std = gain / sqrt(fan_in_effective) instead of the reference's 0.01."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in hifigan_layer_shapes(cfg):
        if name.startswith("ups."):
            cin, cout, k = shape
            stride = cfg.upsample_rates[int(name.split(".")[1])]
            fan = cin * k / stride                       # taps that actually hit one output sample
            bias_n = cout
        else:
            cout, cin, k = shape
            fan = cin * k
            bias_n = cout
        gain = 0.6 if ".convs" in name else 1.0          # residual branches: keep the sum from blowing up
        sd[name + ".weight"] = _randn(g, *shape, std=gain * fan ** -0.5)
        sd[name + ".bias"] = _randn(g, bias_n, std=0.05)
    return sd


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------

def synthetic_logmel(gen: torch.Generator, *shape: int) -> torch.Tensor:
    """Log-mel-like values ~ clip(N(-6.3, 1.95^2), -11.5, 1) (hifi-gan/hifigan_test statistics)."""
    return (torch.randn(*shape, generator=gen) * 1.95 - 6.3).clamp_(-11.5, 1.0)


def synthetic_flow_inputs(cfg: FlowConfig, B: int, N: int, prompt: int = 150, seed: int = 30):
    """ids / cond / y0 / mask shaped as the generation scripts build them
    (monologue_generation.py:160-171, dialogue_generation.py:307-323)."""
    g = torch.Generator().manual_seed(seed)
    prompt = min(prompt, N)
    ids_shape = (B, N, 2) if cfg.n_streams == 2 else (B, N)
    ids = torch.randint(0, 501, ids_shape, generator=g, dtype=torch.int64)
    cond = torch.zeros(B, N, cfg.dim_in)
    cond[:, :prompt] = synthetic_logmel(g, B, prompt, cfg.dim_in)
    y0 = torch.randn(B, N, cfg.dim_x, generator=g)
    mask = torch.zeros(B, N, dtype=torch.bool)
    mask[:, prompt:] = True
    return ids, cond, y0, mask


# --------------------------------------------------------------------------------------
# text-to-semantic (CoSingle / CoMix): SURVEY.md section 8f rank 1
# --------------------------------------------------------------------------------------

@dataclass(frozen=True)
class T2SConfig:
    """Arguments of ``TextToSemantic(...)`` (covomix/covomix_model/text2semantic.py:417-451) as passed by
    conditional_model.py:122-135 with running_command/T2S_Co{Single,Mix}.sh."""
    dim: int = 512                       # --CoVoMix_dim_transformer: encoder width == cross-attention context width
    source_depth: int = 4
    target_depth: int = 4
    heads: int = 8
    dim_head: int = 64
    num_text_token_ids: int = 30530
    num_semantic_token_ids: int = 501    # --text2semantic_tokens
    two_output: bool = False             # --text2semantic_two_output (CoMix)
    target_transformer_dim: int = 512    # --target_transformer_dim (1024 for CoMix)
    ff_mult: int = 4
    text_pad_id: int = 0
    semantic_pad_id: int = -1

    @property
    def inner(self) -> int:              # heads * dim_head (text2semantic.py:190-191)
        return self.heads * self.dim_head

    def ff_inner(self, dim: int) -> int:  # text2semantic.py:161
        return int(dim * self.ff_mult * 2 / 3)

    @property
    def n_out(self) -> int:
        return 2 if self.two_output else 1

    @property
    def dim_emb(self) -> int:            # width of one semantic embedding row (text2semantic.py:512-515)
        return self.target_transformer_dim // self.n_out

    @property
    def text_eos_id(self) -> int:        # text2semantic.py:491-494
        return self.num_text_token_ids

    @property
    def semantic_eos_id(self) -> int:
        return self.num_semantic_token_ids

    @property
    def n_logits(self) -> int:           # tied to the embedding table, which has the EOS row (text2semantic.py:541)
        return self.num_semantic_token_ids + 1


COSINGLE = T2SConfig()
COMIX = T2SConfig(two_output=True, target_transformer_dim=1024)


def synthetic_t2s_state_dict(cfg: T2SConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """State dict of ``TextToSemantic`` (keys as in its ``state_dict()``; the three tied copies of the semantic
    embedding and the two of the text embedding are the same tensor).  Norm gains are perturbed away from 1."""
    g = torch.Generator().manual_seed(seed)
    d, dt, inner = cfg.dim, cfg.target_transformer_dim, cfg.inner
    sd: Dict[str, torch.Tensor] = {}
    # std 0.1: logits = emb . h have std ~ 0.1 * sqrt(dim_emb) ~ 2, so that top-k + Gumbel sampling actually depends
    # on the noise (with N(0,1) rows the tied logits have std ~ 22 and decoding is greedy whatever the seed)
    sem = _randn(g, cfg.n_logits, cfg.dim_emb, std=0.1)
    txt = _randn(g, cfg.num_text_token_ids + 1, d)
    sd["semantic_token_emb.weight"] = sem
    sd["token_emb.speech.weight"] = sem
    sd["token_emb.text.weight"] = txt
    sd["start_token.speech"] = _randn(g, dt)
    sd["start_token.text"] = _randn(g, d)
    sd["to_logits.speech.weight"] = sem
    sd["to_logits.text.weight"] = txt
    inv_freq = 1.0 / (10000 ** (torch.arange(0, cfg.dim_head, 2).float() / cfg.dim_head))

    def attn(p: str, dim: int, dim_ctx: int, cross: bool):
        if not cross:
            sd[p + "rotary_emb.freqs"] = inv_freq.clone()
        else:
            sd[p + "null_kv"] = _randn(g, 2, cfg.heads, 1, cfg.dim_head)
        sd[p + "norm.gamma"] = 1.0 + _randn(g, dim, std=0.1)
        sd[p + "to_q.0.weight"] = _randn(g, inner, dim, std=dim ** -0.5)
        sd[p + "to_kv.0.weight"] = _randn(g, 2 * inner, dim_ctx, std=dim_ctx ** -0.5)
        sd[p + "to_out.weight"] = _randn(g, dim, inner, std=inner ** -0.5)

    def ff(p: str, dim: int):
        fi = cfg.ff_inner(dim)
        sd[p + "0.gamma"] = 1.0 + _randn(g, dim, std=0.1)
        sd[p + "1.weight"] = _randn(g, 2 * fi, dim, std=dim ** -0.5)
        sd[p + "1.bias"] = _randn(g, 2 * fi, std=0.02)
        sd[p + "4.weight"] = _randn(g, dim, fi, std=fi ** -0.5)
        sd[p + "4.bias"] = _randn(g, dim, std=0.02)

    for L in range(cfg.source_depth):
        attn(f"source_transformer.layers.{L}.0.", d, d, False)
        ff(f"source_transformer.layers.{L}.2.", d)
    sd["source_transformer.final_norm.gamma"] = 1.0 + _randn(g, d, std=0.1)
    for L in range(cfg.target_depth):
        attn(f"target_transformer.layers.{L}.0.", dt, dt, False)
        attn(f"target_transformer.layers.{L}.1.", dt, d, True)
        ff(f"target_transformer.layers.{L}.2.", dt)
    sd["target_transformer.final_norm.gamma"] = 1.0 + _randn(g, dt, std=0.1)
    return sd


def synthetic_text_ids(cfg: T2SConfig, B: int, S: int, seed: int = 30, ragged: bool = True) -> torch.Tensor:
    """BERT-tokenizer-like ids in [1, num_text_token_ids), right-padded with ``text_pad_id`` when ragged
    (``tokenizer([txt], padding=True)``, dialogue_generation.py:335)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, cfg.num_text_token_ids, (B, S), generator=g, dtype=torch.int64)
    if ragged:
        for b in range(1, B):
            keep = max(1, S - (b * 3) % max(S // 2, 1))
            ids[b, keep:] = cfg.text_pad_id
    return ids
