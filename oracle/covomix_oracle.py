"""CPU oracle for the CoVoMix inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 PyTorch-on-CPU restatement of the reference's algorithm for the
path named in BASELINE.json (flow-matching acoustic decoder + HiFi-GAN generator).  It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, as the checker / the CPU arm.  The product
(``neurips2024-covomix_b200``) never imports it and has no CPU fallback.

Parity status: PINNED against the reference itself.  ``tests/golden/make_golden.py`` imports
the real reference modules from ``/root/reference`` (``covomix/covomix_model/acoustic.py`` and
``hifi-gan/models.py``; stubs only for the absent ``matplotlib`` / ``torchode`` imports and for
``torchdiffeq``), loads the same seeded state dicts into them, and writes their outputs to
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against those vectors.
The one piece that cannot be pinned against reference-owned code is the ODE solver: the
reference calls the third-party ``torchdiffeq.odeint`` (acoustic.py:12,656; version unpinned,
transitive dependency of voicebox_pytorch==0.0.34, README.md:47), which is absent here.
``odeint_fixed_grid`` restates torchdiffeq's published fixed-grid algorithm
(``FixedGridODESolver.integrate`` + ``Midpoint._step_func`` / ``Euler._step_func``) and is
checked by a closed-form known-answer test (linear ODE) -- "parity unpinned" for that
function only.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LRELU_SLOPE = 0.1                     # hifi-gan/models.py:8


# ======================================================================================
# ODE solver (third-party torchdiffeq, restated)
# ======================================================================================

def fixed_grid(t0: float, t1: float, step_size: float) -> Tensor:
    """torchdiffeq ``_grid_constructor_from_step_size``: ``arange(0, ceil((t1-t0)/h + 1))*h + t0``
    with the last point snapped to ``t1``.  Call site: acoustic.py:586-591 (step_size 0.0625)."""
    niters = int(math.ceil((t1 - t0) / step_size + 1))
    grid = torch.arange(0, niters, dtype=torch.float32) * step_size + t0
    grid[-1] = t1
    return grid


def odeint_fixed_grid(fn: Callable[[Tensor, Tensor], Tensor], y0: Tensor, t: Tensor,
                      method: str = "midpoint", step_size: float = 0.0625) -> Tensor:
    """torchdiffeq ``FixedGridODESolver.integrate`` for ``method in {'euler','midpoint'}``.

    Midpoint step: ``f0 = f(t0, y); dy = dt * f(t0 + dt/2, y + f0*dt/2)``; Euler step:
    ``dy = dt * f(t0, y)``.  Outputs at the requested times ``t`` by linear interpolation inside
    the step that contains them (exact hits return the grid value).  Reference call:
    ``odeint(fn, y0, t, atol, rtol, method='midpoint', options={'step_size': 0.0625})``
    (acoustic.py:656 with kwargs :586-591; atol/rtol are ignored by fixed-grid solvers)."""
    grid = fixed_grid(float(t[0]), float(t[-1]), step_size)
    sol = [y0]
    j = 1
    y = y0
    for k in range(len(grid) - 1):
        ta, tb = grid[k], grid[k + 1]
        dt = tb - ta
        if method == "midpoint":
            f0 = fn(ta, y)
            dy = dt * fn(ta + 0.5 * dt, y + f0 * (0.5 * dt))
        elif method == "euler":
            dy = dt * fn(ta, y)
        else:
            raise ValueError(method)
        y1 = y + dy
        while j < len(t) and tb >= t[j]:
            if t[j] == ta:
                sol.append(y)
            elif t[j] == tb:
                sol.append(y1)
            else:
                sol.append(y + (t[j] - ta) / (tb - ta) * (y1 - y))
            j += 1
        y = y1
    return torch.stack(sol)


def ode_eval_times(method: str, n_steps: int) -> Sequence[float]:
    """The distinct times at which the velocity net is evaluated."""
    h = 1.0 / n_steps
    ts = []
    for k in range(n_steps):
        ts.append(k * h)
        if method == "midpoint":
            ts.append(k * h + 0.5 * h)
    return ts


# ======================================================================================
# velocity net (covomix/covomix_model/acoustic.py + attend.py)
# ======================================================================================

def rms_normalize(x: Tensor, dim_scale: float) -> Tensor:
    """``F.normalize(x, dim=-1) * sqrt(dim)`` (acoustic.py:175,199): x / max(||x||_2, 1e-12)."""
    return F.normalize(x, dim=-1) * dim_scale


def time_embedding(sd: Dict[str, Tensor], times: Tensor) -> Tensor:
    """``sinu_pos_emb`` Sequential (acoustic.py:361-365; LearnedSinusoidalPosEmb :107-111):
    [sin(2*pi*t*w) | cos(2*pi*t*w)] -> Linear -> SiLU.  times: [B] -> [B, 4*dim]."""
    freqs = times[:, None] * sd["sinu_pos_emb.0.weights"][None, :] * 2 * math.pi
    four = torch.cat((freqs.sin(), freqs.cos()), dim=-1)
    return F.silu(F.linear(four, sd["sinu_pos_emb.1.weight"], sd["sinu_pos_emb.1.bias"]))


def rotary_table(sd: Dict[str, Tensor], n: int) -> Tensor:
    """RotaryEmbedding.forward (acoustic.py:126-130): freqs[n, j] duplicated to dim_head."""
    inv_freq = sd["transformer.rotary_emb.inv_freq"]
    t = torch.arange(n, dtype=torch.float32, device=inv_freq.device)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    return torch.cat((freqs, freqs), dim=-1)


def apply_rotary(pos: Tensor, t: Tensor) -> Tensor:
    """acoustic.py:132-137: t*cos + rotate_half(t)*sin, rotate_half = cat(-x2, x1) (NeoX halves)."""
    x1, x2 = t.chunk(2, dim=-1)
    return t * pos.cos() + torch.cat((-x2, x1), dim=-1) * pos.sin()


def adaptive_rmsnorm(sd: Dict[str, Tensor], prefix: str, x: Tensor, time_emb: Tensor) -> Tensor:
    """AdaptiveRMSNorm.forward (acoustic.py:198-204)."""
    normed = rms_normalize(x, x.shape[-1] ** 0.5)
    gamma = F.linear(time_emb, sd[prefix + ".to_gamma.weight"], sd[prefix + ".to_gamma.bias"])
    beta = F.linear(time_emb, sd[prefix + ".to_beta.weight"], sd[prefix + ".to_beta.bias"])
    return normed * gamma[:, None, :] + beta[:, None, :]


def attention(sd: Dict[str, Tensor], prefix: str, x: Tensor, rotary: Tensor, heads: int) -> Tensor:
    """Attention.forward (acoustic.py:225-237) with Attend.forward's non-flash branch
    (attend.py:110-124): sim = q k^T * dh^-0.5, softmax, attn v.  No mask at inference."""
    B, N, _ = x.shape
    qkv = F.linear(x, sd[prefix + ".to_qkv.weight"])
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (t.reshape(B, N, heads, -1).permute(0, 2, 1, 3) for t in (q, k, v))
    q, k = apply_rotary(rotary, q), apply_rotary(rotary, k)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (q.shape[-1] ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn.to(v.dtype), v)      # (.to: no-op in fp32; lets tests run this under autocast)
    out = out.permute(0, 2, 1, 3).reshape(B, N, -1)
    return F.linear(out, sd[prefix + ".to_out.weight"])


def transformer(sd: Dict[str, Tensor], x: Tensor, time_emb: Tensor, depth: int, heads: int) -> Tensor:
    """Transformer.forward (acoustic.py:288-318): U-Net skips (first half push their input,
    second half pop + Linear(2d->d)), pre-norm attention and feed-forward, final RMSNorm."""
    rotary = rotary_table(sd, x.shape[-2])
    skips = []
    for L in range(depth):
        p = f"transformer.layers.{L}"
        if L + 1 <= depth // 2:
            skips.append(x)
        else:
            x = torch.cat((x, skips.pop()), dim=-1)
            x = F.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
        a_in = adaptive_rmsnorm(sd, p + ".1", x, time_emb)
        x = attention(sd, p + ".2", a_in, rotary, heads) + x
        f_in = adaptive_rmsnorm(sd, p + ".3", x, time_emb)
        h = F.gelu(F.linear(f_in, sd[p + ".4.0.weight"], sd[p + ".4.0.bias"]))       # FeedForward :241-246
        x = F.linear(h, sd[p + ".4.2.weight"], sd[p + ".4.2.bias"]) + x
    return rms_normalize(x, x.shape[-1] ** 0.5) * sd["transformer.final_norm.gamma"]  # RMSNorm :174-175


def velocity(sd: Dict[str, Tensor], cfg, x: Tensor, ids: Tensor, cond: Tensor, times: Tensor,
             drop_cond: bool) -> Tensor:
    """CoVoMix.forward, inference part (acoustic.py:430-521) for cond_drop_prob in {0, 1}."""
    B = cond.shape[0]
    if times.ndim == 0:                                                   # :452-456
        times = times.repeat(B)
    if drop_cond:                                                         # :473-494
        cond = sd["null_cond"].expand_as(cond)
        ids = torch.full_like(ids, cfg.num_phoneme_tokens)
    emb = F.embedding(ids, sd["to_phoneme_emb.weight"])                  # :496
    if emb.ndim == 4:                                                     # :499-500
        emb = emb.reshape(emb.shape[0], emb.shape[1], 2 * cfg.dim_phoneme_emb)
    h = F.linear(torch.cat((x, emb, cond), dim=-1), sd["to_embed.weight"], sd["to_embed.bias"])  # :503-505
    conv = F.conv1d(h.transpose(1, 2), sd["conv_embed.dw_conv1d.0.weight"], sd["conv_embed.dw_conv1d.0.bias"],
                    padding=cfg.conv_pos_kernel // 2, groups=cfg.dim)     # :153-161
    h = F.gelu(conv).transpose(1, 2) + h                                  # :508
    te = time_embedding(sd, times)                                        # :510
    h = transformer(sd, h, te, cfg.depth, cfg.heads)                      # :514
    return F.linear(h, sd["to_pred.weight"])                              # :516


def velocity_cfg(sd, cfg, x, ids, cond, times, cond_scale: float) -> Tensor:
    """CoVoMix.forward_with_cond_scale (acoustic.py:414-428): v*(1+s) - s*v_null."""
    v = velocity(sd, cfg, x, ids, cond, times, drop_cond=False)
    if cond_scale == 1.0:
        return v
    vn = velocity(sd, cfg, x, ids, cond, times, drop_cond=True)
    return v * (1 + cond_scale) - cond_scale * vn


@torch.inference_mode()
def flow_sample(sd, cfg, ids: Tensor, cond: Tensor, y0: Tensor, cond_scale: float = 1.0,
                method: str = "midpoint", step_size: float = 0.0625, steps: int = 3) -> Tensor:
    """ConditionalFlowMatcherWrapper.sample (acoustic.py:597-688) with y0 supplied by the
    caller (the reference draws it with torch.randn_like at :647-650)."""
    t = torch.linspace(0, 1, steps)
    fn = lambda tt, xx: velocity_cfg(sd, cfg, xx, ids, cond, tt, cond_scale)
    return odeint_fixed_grid(fn, y0, t, method=method, step_size=step_size)[-1]


# ======================================================================================
# HiFi-GAN generator (hifi-gan/models.py == covomix/vocoder/models.py)
# ======================================================================================

def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """hifi-gan/utils.py:34-35."""
    return int((kernel_size * dilation - dilation) / 2)


def resblock1(sd, prefix: str, x: Tensor, k: int, dils: Sequence[int]) -> Tensor:
    """ResBlock1.forward (models.py:35-42)."""
    for m, d in enumerate(dils):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{prefix}.convs1.{m}.weight"], sd[f"{prefix}.convs1.{m}.bias"],
                      dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{prefix}.convs2.{m}.weight"], sd[f"{prefix}.convs2.{m}.bias"],
                      padding=get_padding(k, 1))
        x = xt + x
    return x


def resblock2(sd, prefix: str, x: Tensor, k: int, dils: Sequence[int]) -> Tensor:
    """ResBlock2.forward (models.py:63-68)."""
    for m, d in enumerate(dils):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{prefix}.convs.{m}.weight"], sd[f"{prefix}.convs.{m}.bias"],
                      dilation=d, padding=get_padding(k, d))
        x = xt + x
    return x


@torch.inference_mode()
def hifigan_forward(sd, cfg, mel: Tensor) -> Tensor:
    """Generator.forward (models.py:100-116) on weights after remove_weight_norm.
    mel [B,80,T] -> [B, 1, hop*T + 32]; unbatched [80,T] -> [1, hop*T + 32] (Conv1d's unbatched convention, which is
    what the reference module returns) for config_covomix.json."""
    unbatched = mel.ndim == 2
    x = mel[None] if unbatched else mel
    x = F.conv1d(x, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    nk = len(cfg.resblock_kernel_sizes)
    rb = resblock1 if cfg.resblock == "1" else resblock2
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            y = rb(sd, f"resblocks.{i * nk + j}", x, cfg.resblock_kernel_sizes[j], cfg.resblock_dilation_sizes[j])
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                                   # default slope 0.01 (models.py:112)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    x = torch.tanh(x)
    return x[0] if unbatched else x


def wav_to_int16(wav: Tensor):
    """mel_decode_to_wav (monologue_generation.py:52-59): squeeze, *32768, numpy astype(int16)."""
    return (wav.squeeze() * 32768.0).cpu().numpy().astype("int16")
