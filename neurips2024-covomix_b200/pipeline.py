"""Batching / post-processing glue around the two accelerated call sites.

The reference scripts loop one utterance at a time (dialogue_generation.py:283): build ``phone_input`` / ``mel_input`` /
``mask`` -> ``model.synthesis_sample`` -> ``sampled[:, mask]`` -> ``mel_decode_to_wav`` -> int16
(monologue_generation.py:160-177).  ``synthesize`` does the same for a LIST of utterances, but groups them into
equal-length batches (the velocity net has no padding mask, so a batch is exact only for equal-length items) for the
sampler and into equal-length batches of generated frames for the vocoder, and returns the int16 waveforms in input order.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from .sharding import plan_batches


def synthesize(sampler, generator, items: Sequence[Dict[str, torch.Tensor]], cond_scale: float = 0.7, batch: int = 8,
               vocoder_batch: int = 8) -> List[np.ndarray]:
    """items[i] = {"phoneme_ids": Long[N] or [N,2], "cond": Float[N,dim_in], "mask": Bool[N][, "y0": Float[N,80]]} (one
    utterance, as the scripts build them) -> list of int16 arrays (``mel_decode_to_wav`` semantics), same order as ``items``."""
    device = sampler.device
    mels: List[torch.Tensor] = [None] * len(items)
    for n, idx in plan_batches([int(it["cond"].shape[0]) for it in items], batch):
        ids = torch.stack([items[i]["phoneme_ids"] for i in idx]).to(device)
        cond = torch.stack([items[i]["cond"] for i in idx]).to(device)
        mask = torch.stack([items[i]["mask"] for i in idx]).to(device)
        kw = {}
        if all("y0" in items[i] for i in idx):                                    # optional: caller-supplied noise (tests)
            kw["y0"] = torch.stack([items[i]["y0"] for i in idx]).to(device)
        sampled = sampler.sample(phoneme_ids=ids, cond=cond, mask=mask, cond_scale=cond_scale, **kw)
        for k, i in enumerate(idx):
            mels[i] = sampled[k][mask[k]].transpose(0, 1).contiguous()          # [80, T_i]: sampled[:, mask].permute(0,2,1)
    out: List[np.ndarray] = [None] * len(items)
    for t, idx in plan_batches([int(m.shape[1]) for m in mels], vocoder_batch):
        if t == 0:
            for i in idx:
                out[i] = np.zeros(0, dtype=np.int16)
            continue
        wav = generator(torch.stack([mels[i] for i in idx]), out_dtype="i16")   # fused x32768 + int16 cast
        wav = wav.reshape(len(idx), -1).cpu().numpy()
        for k, i in enumerate(idx):
            out[i] = wav[k]
    return out


def concat_turns(turns: Sequence[np.ndarray]) -> np.ndarray:
    """dialogue_generation.py:189: per-turn audio of `--mode covosingle` dialogues is concatenated."""
    return np.concatenate(list(turns)) if len(turns) else np.zeros(0, dtype=np.int16)


def comix_pred(t2s, text_ids: torch.Tensor, **kw):
    """``comix_pred`` (dialogue_generation.py:332-344 == monologue_generation.py:307-319) on top of
    ``B200TextToSemantic.sample``: one tokenised text -> the two semantic streams (first / second half of the flattened
    output, exactly the reference's split) and the empty ``mel_to_synthesis`` placeholder."""
    semantic = t2s.sample(text_ids.reshape(1, -1), **kw).reshape(-1).cpu()
    half = semantic.shape[0] // 2
    s1, s2 = semantic[:half], semantic[half:]
    return s1, s2, torch.zeros((80, len(s1)))


def dialogue_item(semantic_a: torch.Tensor, semantic_b: torch.Tensor, mel_prompt: torch.Tensor, new_a: torch.Tensor,
                  new_b: torch.Tensor, pad_id: int = 157, max_id: int = 501) -> Dict[str, torch.Tensor]:
    """The id / cond / mask assembly of ``covomix(...)`` (dialogue_generation.py:307-321): prompt streams followed by the
    generated ones, right-padded with id 157 to a common length, clamped to 501; ``cond`` carries the prompt mel and zeros;
    ``mask`` selects the generated frames."""
    n_prompt = mel_prompt.shape[0]
    a = torch.cat((semantic_a[:n_prompt], new_a))
    b = torch.cat((semantic_b[:n_prompt], new_b))
    n = max(a.shape[0], b.shape[0])
    a = torch.nn.functional.pad(a, (0, n - a.shape[0]), value=pad_id)
    b = torch.nn.functional.pad(b, (0, n - b.shape[0]), value=pad_id)
    ids = torch.stack((a, b), dim=-1).clamp(max=max_id)
    mask = torch.zeros(n, dtype=torch.bool)
    mask[n_prompt:] = True
    cond = torch.zeros(n, mel_prompt.shape[1])
    cond[:n_prompt] = mel_prompt
    return {"phoneme_ids": ids, "cond": cond, "mask": mask}


def covomix_dialogues(t2s, sampler, generator, texts: Sequence[torch.Tensor], prompts: Sequence[Dict[str, torch.Tensor]],
                      cond_scale: float = 0.7, batch: int = 8, t2s_kwargs: Dict = None) -> List[np.ndarray]:
    """The ``covomix(...)`` loop of dialogue_generation.py:283-330 for a LIST of dialogues: per dialogue ``comix_pred`` on the
    tokenised text (text-to-semantic, one call each -- its EOS rule couples the rows of a batch, so dialogues are decoded
    one at a time like the reference), ``dialogue_item`` to assemble ids / cond / mask behind the prompt, then ``synthesize``
    batches equal-length items through the acoustic model and the vocoder.

    texts[i]: Long[S_i] BERT-tokenised text; prompts[i] = {"semantic_a": Long[P], "semantic_b": Long[P], "mel": Float[P, 160]}
    (``prepare_oracle_hubert`` output for both speakers, cut to the common length and concatenated, :287-297)."""
    items = []
    for text, pr in zip(texts, prompts):
        s1, s2, _ = comix_pred(t2s, text, **(t2s_kwargs or {}))
        items.append(dialogue_item(pr["semantic_a"], pr["semantic_b"], pr["mel"], s1, s2))
    return synthesize(sampler, generator, items, cond_scale=cond_scale, batch=batch)
