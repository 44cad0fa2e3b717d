"""Generate tests/golden/t2s_*.npz by running the REAL reference ``TextToSemantic`` (this container only).

    python tests/golden/make_golden_t2s.py

Imports ``covomix/covomix_model/text2semantic.py`` read-only from /root/reference (it imports cleanly here: beartype,
einops, transformers are present), builds the model exactly as ``covomix/conditional_model.py:122-135`` does for
running_command/T2S_CoMix.sh / T2S_CoSingle.sh, loads the seeded state dict of ``synthetic.synthetic_t2s_state_dict``
and records: the encoder output, the logits of every decoding step (forward hook on ``to_logits['speech']``), the
sampled tokens / mask, and the uniform noise the run consumed (replayed from the same seed: ``gumbel_noise`` draws
``zeros_like(logits).uniform_(0, 1)`` once per stream per step and nothing else touches the RNG in eval mode).
The GPU box has no /root/reference: only the .npz files travel.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import covomix_b200  # noqa: E402,F401
from covomix_b200 import synthetic as syn  # noqa: E402


def ref_model(cfg: syn.T2SConfig, sd):
    from covomix.covomix_model.text2semantic import TextToSemantic
    m = TextToSemantic(dim=cfg.dim, source_depth=cfg.source_depth, target_depth=cfg.target_depth, semantic_pad_id=-1,
                       text_pad_id=0, heads=cfg.heads, num_text_token_ids=cfg.num_text_token_ids,
                       num_semantic_token_ids=cfg.num_semantic_token_ids, no_source_transformer=False,
                       two_output=cfg.two_output, two_input=False, target_transformer_dim=cfg.target_transformer_dim)
    m.load_state_dict(sd, strict=True)
    return m.eval()


_MODELS = {}


def run_case(name, cfg, B, S, max_length, seed, input_seed=30, ragged=True):
    if cfg not in _MODELS:
        _MODELS[cfg] = ref_model(cfg, syn.synthetic_t2s_state_dict(cfg, seed=1234))
    m = _MODELS[cfg]
    ids = syn.synthetic_text_ids(cfg, B, S, seed=input_seed, ragged=ragged)
    logits, enc = [], []
    h1 = m.to_logits["speech"].register_forward_hook(lambda mod, inp, out: logits.append(out[:, -1].clone()))
    h2 = m.source_transformer.register_forward_hook(lambda mod, inp, out: enc.append(out.clone()))
    torch.manual_seed(seed)
    target, mask = m.generate(ids.clone(), source_type="text", target_type="speech", return_target_mask=True,
                              max_length=max_length)
    h1.remove()
    h2.remove()
    n_out = cfg.n_out
    steps = len(logits) // n_out
    torch.manual_seed(seed)
    u = torch.stack([torch.stack([torch.zeros(B, cfg.n_logits).uniform_(0, 1) for _ in range(n_out)])
                     for _ in range(steps)])
    lg = torch.stack(logits).view(steps, n_out, B, cfg.n_logits)
    print(name, "steps", steps, "target", tuple(target.shape), "kept", int(mask.sum()), "eos", int((target == 501).sum()))
    np.savez_compressed(os.path.join(HERE, f"t2s_{name}.npz"), B=B, S=S, max_length=max_length, weight_seed=1234,
                        input_seed=input_seed, noise_seed=seed, ragged=ragged, steps=steps,
                        enc=enc[0].numpy(), logits=lg.numpy(), target=target.numpy(), mask=mask.numpy(),
                        u=u.numpy().astype(np.float32))
    return steps


@torch.inference_mode()
def main():
    # CoMix (two output streams, target dim 1024), ragged batch of 2: masks on both attention types
    run_case("comix_b2", syn.COMIX, B=2, S=12, max_length=20, seed=77)
    # CoMix, batch 1 (how the scripts call it), run until an EOS ends the loop early
    for seed in range(100, 200):
        sd_steps = run_case("comix_eos", syn.COMIX, B=1, S=9, max_length=48, seed=seed, ragged=False)
        if sd_steps < 40:
            break
    # CoSingle (one stream, target dim 512)
    run_case("cosingle_b1", syn.COSINGLE, B=1, S=10, max_length=24, seed=78, ragged=False)


if __name__ == "__main__":
    main()
