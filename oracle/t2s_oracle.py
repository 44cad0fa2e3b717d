"""CPU oracle for the CoSingle / CoMix text-to-semantic decoder.
TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs;
the product (``neurips2024-covomix_b200``) never imports it and has no CPU fallback.

Plain fp32 PyTorch-on-CPU restatement of ``TextToSemantic.generate`` (the non-beam, non-speculative branch the
generation scripts use: ``TextToSemanticWrapper.sample`` -> ``generate(source_type='text', target_type='speech')``,
covomix/covomix_model/text2semantic.py:1237-1251, :659-848) over a state dict in the reference's key layout.

Parity status: PINNED against the reference itself.  ``tests/golden/make_golden_t2s.py`` imports the real
``covomix/covomix_model/text2semantic.py`` from /root/reference, loads the seeded state dicts of
``synthetic.synthetic_t2s_state_dict`` into ``TextToSemantic`` and records encoder output, per-step logits, sampled tokens
and the replayed uniform noise; ``tests/test_t2s_oracle.py`` holds this file to those vectors (logits rel-L2 < 1e-5,
tokens identical).

The only deliberate difference: the reference draws its Gumbel noise from the global torch RNG inside the loop
(``gumbel_noise``, text2semantic.py:108-113); here the uniform draws are an explicit input ``u[step, stream, B, n_logits]``
so that the CUDA path and the oracle can consume the same numbers.  Every function cites the reference file:line it
follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def rmsnorm(x: Tensor, gamma: Tensor) -> Tensor:
    """``RMSNorm.forward`` text2semantic.py:143-151: ``F.normalize(x, dim=-1) * sqrt(dim) * gamma``."""
    return F.normalize(x, dim=-1) * (x.shape[-1] ** 0.5) * gamma


def rotary(t: Tensor, inv_freq: Tensor, positions: Tensor) -> Tensor:
    """``RotaryEmbedding.rotate_queries_or_keys`` + ``apply_rotary_emb`` + ``rotate_half``
    (covomix/covomix_model/rotary_embedding_torch.py:132-144, :43-52, :36-40): INTERLEAVED pairs (2i, 2i+1), angle
    ``pos * inv_freq[i]`` for both members of the pair.  t: [..., n, d]; positions: [n]."""
    freqs = positions.to(t.dtype)[:, None] * inv_freq[None, :]            # [n, d/2]
    freqs = freqs.repeat_interleave(2, dim=-1)                            # '... n -> ... (n r)', r = 2
    x1, x2 = t[..., 0::2], t[..., 1::2]
    rot = torch.stack((-x2, x1), dim=-1).flatten(-2)
    return t * freqs.cos() + rot * freqs.sin()


def attend(q: Tensor, k: Tensor, v: Tensor, key_mask: Optional[Tensor], causal: bool) -> Tensor:
    """``Attend.forward`` (non-flash branch) covomix/covomix_model/attend_t2s.py:127-171.
    q [B,H,i,d], k/v [B,H,j,d], key_mask bool [B,j] (True = keep)."""
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (q.shape[-1] ** -0.5)
    if key_mask is not None:
        sim = sim.masked_fill(~key_mask[:, None, None, :], -torch.finfo(sim.dtype).max)
    if causal:
        i, j = sim.shape[-2:]
        n = max(i, j)
        cm = torch.ones(n, n, dtype=torch.bool).triu(1)[-i:, :]
        sim = sim.masked_fill(cm, -torch.finfo(sim.dtype).max)
    return torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)


def _heads(t: Tensor, h: int) -> Tensor:
    b, n, _ = t.shape
    return t.view(b, n, h, -1).transpose(1, 2)                             # 'b n (h d) -> b h n d'


def feedforward(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """``FeedForward`` + ``GEGLU`` text2semantic.py:155-168: RMSNorm -> Linear(dim, 2*inner) -> gelu(gate) * x
    (x = first half, gate = second half) -> Linear(inner, dim)."""
    h = F.linear(rmsnorm(x, sd[p + "0.gamma"]), sd[p + "1.weight"], sd[p + "1.bias"])
    a, gate = h.chunk(2, dim=-1)
    return F.linear(F.gelu(gate) * a, sd[p + "4.weight"], sd[p + "4.bias"])


def encode(sd: Dict[str, Tensor], cfg, text_ids: Tensor) -> Tuple[Tensor, Tensor]:
    """Source side of ``generate`` (text2semantic.py:716-744): ``set_eos_id`` (:57-66), mask = ids != pad, embedding,
    ``source_transformer`` (``Transformer.forward`` :308-375 with causal=False, no cross attention).
    Returns (source_emb [B,S+1,dim], source_mask [B,S+1])."""
    ids = set_eos_id(text_ids, cfg.text_eos_id, cfg.text_pad_id)
    mask = ids != cfg.text_pad_id
    x = sd["token_emb.text.weight"][ids]
    pos = torch.arange(ids.shape[1])
    for L in range(cfg.source_depth):
        p = f"source_transformer.layers.{L}."
        h = rmsnorm(x, sd[p + "0.norm.gamma"])                              # Attention.forward :224-270
        q = _heads(F.linear(h, sd[p + "0.to_q.0.weight"]), cfg.heads)
        k, v = F.linear(h, sd[p + "0.to_kv.0.weight"]).chunk(2, dim=-1)     # '(kv h d)': k = first half
        k, v = _heads(k, cfg.heads), _heads(v, cfg.heads)
        inv = sd[p + "0.rotary_emb.freqs"]
        q, k = rotary(q, inv, pos), rotary(k, inv, pos)
        o = attend(q, k, v, mask, causal=False).transpose(1, 2).flatten(2)
        x = F.linear(o, sd[p + "0.to_out.weight"]) + x
        x = feedforward(sd, p + "2.", x) + x
    return rmsnorm(x, sd["source_transformer.final_norm.gamma"]), mask


def set_eos_id(t: Tensor, eos_id: int, pad_id: int) -> Tensor:
    """text2semantic.py:57-66: append one pad column, then write EOS at the first pad position of each row."""
    eos_idx = ((t == pad_id).cumsum(dim=-1) == 0).sum(dim=-1, keepdim=True).long()
    t = F.pad(t, (0, 1), value=pad_id)
    t[torch.arange(t.shape[0])[:, None], eos_idx] = eos_id
    return t


def mask_after_eos(target: Tensor, eos_id: int, pad_id: int) -> Tensor:
    """text2semantic.py:72-75."""
    m = (target == eos_id).cumsum(dim=-1) > 0
    m = F.pad(m, (1, -1), value=False)
    return target.masked_fill(m, pad_id)


class DecoderState:
    """KV caches of the target transformer (``Transformer.forward(cache=...)`` text2semantic.py:326-345 keeps
    un-rotated self-attention k/v per layer; the cross-attention k/v of the fixed context are recomputed by the
    reference every step (:231) and computed once here -- same values)."""

    def __init__(self, sd, cfg, source_emb: Tensor, source_mask: Tensor):
        self.sd, self.cfg = sd, cfg
        self.self_k: List[Optional[Tensor]] = [None] * cfg.target_depth
        self.self_v: List[Optional[Tensor]] = [None] * cfg.target_depth
        self.ctx_k, self.ctx_v = [], []
        b = source_emb.shape[0]
        for L in range(cfg.target_depth):
            p = f"target_transformer.layers.{L}.1."
            k, v = F.linear(source_emb, sd[p + "to_kv.0.weight"]).chunk(2, dim=-1)
            nk, nv = sd[p + "null_kv"][0], sd[p + "null_kv"][1]                # [H,1,d]; prepended (:253-257)
            self.ctx_k.append(torch.cat((nk.expand(b, -1, -1, -1), _heads(k, cfg.heads)), dim=-2))
            self.ctx_v.append(torch.cat((nv.expand(b, -1, -1, -1), _heads(v, cfg.heads)), dim=-2))
        self.ctx_mask = F.pad(source_mask, (1, 0), value=True)               # :259-260
        self.pos = 0

    def step(self, x: Tensor) -> Tensor:
        """One new position through the target transformer + final norm.  x [B,1,Dt] -> [B,1,Dt]."""
        sd, cfg = self.sd, self.cfg
        for L in range(cfg.target_depth):
            p = f"target_transformer.layers.{L}."
            h = rmsnorm(x, sd[p + "0.norm.gamma"])
            q = _heads(F.linear(h, sd[p + "0.to_q.0.weight"]), cfg.heads)
            k, v = F.linear(h, sd[p + "0.to_kv.0.weight"]).chunk(2, dim=-1)
            k, v = _heads(k, cfg.heads), _heads(v, cfg.heads)
            if self.self_k[L] is not None:
                k = torch.cat((self.self_k[L], k), dim=-2)
                v = torch.cat((self.self_v[L], v), dim=-2)
            self.self_k[L], self.self_v[L] = k, v
            inv = sd[p + "0.rotary_emb.freqs"]
            klen = k.shape[-2]
            # rotate_queries_with_cached_keys (rotary_embedding_torch.py:146-157): q sits at position klen-1
            qr = rotary(q, inv, torch.arange(klen - 1, klen))
            kr = rotary(k, inv, torch.arange(klen))
            o = attend(qr, kr, v, None, causal=True).transpose(1, 2).flatten(2)
            x = F.linear(o, sd[p + "0.to_out.weight"]) + x
            h = rmsnorm(x, sd[p + "1.norm.gamma"])
            q = _heads(F.linear(h, sd[p + "1.to_q.0.weight"]), cfg.heads)
            o = attend(q, self.ctx_k[L], self.ctx_v[L], self.ctx_mask, causal=False).transpose(1, 2).flatten(2)
            x = F.linear(o, sd[p + "1.to_out.weight"]) + x
            x = feedforward(sd, p + "2.", x) + x
        self.pos += 1
        return rmsnorm(x, sd["target_transformer.final_norm.gamma"])


def top_k_filter(logits: Tensor, thres: float = 0.1) -> Tensor:
    """``top_k`` text2semantic.py:126-132: keep the ceil(thres * n) largest, the rest -> -inf."""
    k = math.ceil(thres * logits.shape[-1])
    val, ind = torch.topk(logits, k, dim=-1)
    out = torch.full_like(logits, float("-inf"))
    out.scatter_(-1, ind, val)
    return out


def gumbel_argmax(logits: Tensor, u: Tensor, temperature: float = 1.0) -> Tensor:
    """``gumbel_sample`` text2semantic.py:104-113 with the uniform draw ``u`` made explicit:
    ``argmax(logits / max(T, 1e-10) - log(-log(u)))`` with ``log(t) = torch.log(t.clamp(min=1e-20))``."""
    lg = lambda t: torch.log(t.clamp(min=1e-20))
    return (logits / max(temperature, 1e-10) + (-lg(-lg(u)))).argmax(dim=-1)


@torch.inference_mode()
def generate(sd: Dict[str, Tensor], cfg, text_ids: Tensor, u: Tensor, max_length: int = 2048, temperature: float = 1.0,
             filter_thres: float = 0.1, forced: Optional[Tensor] = None, return_logits: bool = False):
    """``TextToSemantic.generate`` text2semantic.py:659-848 (source_type='text', target_type='speech', cond_scale=1,
    no beam / speculative decoding).

    u: uniform noise [>=steps, n_out, B, n_logits] consumed in the reference's draw order (stream 1 then stream 2 per
    step, :793-800).  forced: optional int64 [B, n_out, L]; when given, the token fed back at step i is forced[..., i]
    (teacher forcing for parity tests; the sampled token is still what is recorded).  Returns
    (target [B, n_out * steps] with -1 after EOS, target_mask, steps[, logits [steps, n_out, B, n_logits]])."""
    source_emb, source_mask = encode(sd, cfg, text_ids)
    B = text_ids.shape[0]
    st = DecoderState(sd, cfg, source_emb, source_mask)
    emb = sd["token_emb.speech.weight"]
    eos, pad = cfg.semantic_eos_id, cfg.semantic_pad_id
    targets = [torch.empty(B, 0, dtype=torch.long) for _ in range(cfg.n_out)]
    x = sd["start_token.speech"].expand(B, 1, -1)
    all_logits = []
    steps = 0
    for i in range(max_length):
        h = st.step(x)[:, -1]                                               # [B, Dt]
        halves = h.chunk(cfg.n_out, dim=-1)                                 # :768-776
        step_logits = []
        for s in range(cfg.n_out):
            logits = F.linear(halves[s], sd["to_logits.speech.weight"])
            step_logits.append(logits)
            sampled = gumbel_argmax(top_k_filter(logits, filter_thres), u[i, s], temperature)
            targets[s] = torch.cat((targets[s], sampled[:, None]), dim=1)
        if return_logits:
            all_logits.append(torch.stack(step_logits))
        steps = i + 1
        fed = [forced[:, s, i] if forced is not None else targets[s][:, -1] for s in range(cfg.n_out)]
        x = torch.cat([emb[f] for f in fed], dim=-1)[:, None, :]           # :746-751
        # EOS logic :804-826
        all_eos = [bool((t == eos).any(dim=-1).all()) for t in targets]
        if not cfg.two_output:
            if not all_eos[0]:
                continue
            targets[0] = mask_after_eos(targets[0], eos, pad)
            break
        targets[0] = mask_after_eos(targets[0], eos, pad)
        if (not all_eos[1]) and (not all_eos[0]):
            continue
        targets[1] = mask_after_eos(targets[1], eos, pad)
        break
    target = torch.cat(targets, dim=1)                                      # :831-832
    mask = target != pad
    if return_logits:
        return target, mask, steps, torch.stack(all_logits)
    return target, mask, steps


def sample(sd, cfg, text_ids: Tensor, u: Tensor, **kw) -> Tensor:
    """``TextToSemanticWrapper.sample`` text2semantic.py:1237-1251: the non-masked part of the target, flattened."""
    target, mask, _ = generate(sd, cfg, text_ids, u, **kw)
    return target[mask]
