"""Text-to-semantic decode timing on a B200: ms per decoding step, tokens/s and the rate at which the decoder matrices
are streamed (the quantity a step is bound by).  python tools/t2s_bench.py [comix|cosingle] [steps]"""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covomix_b200  # noqa: E402,F401
from covomix_b200 import synthetic as syn  # noqa: E402
from covomix_b200.t2s import B200TextToSemantic  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "comix"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    cfg = syn.COMIX if name == "comix" else syn.COSINGLE
    dev = torch.device("cuda:0")
    sd = syn.synthetic_t2s_state_dict(cfg, 1234)
    fmts = os.environ.get("T2S_FMTS", "bf16,fp32").split(",")
    batches = [int(b) for b in os.environ.get("T2S_B", "1,2,4,8").split(",")]
    for fmt in fmts:
        m = B200TextToSemantic(sd, cfg, dev, weight_format=fmt, sm_limit=int(os.environ.get("T2S_SMS", "0")) or None)
        for B in batches:
            ids = syn.synthetic_text_ids(cfg, B, 200, seed=4, ragged=False).to(dev)
            u = torch.rand(steps, cfg.n_out, B, cfg.n_logits, device=dev)
            for _ in range(2):
                m.generate(ids, max_length=steps, noise=u, ignore_eos=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                m.generate(ids, max_length=steps, noise=u, ignore_eos=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            us_step = ms * 1e3 / steps
            gbs = m.weight_bytes_per_step() / (us_step * 1e-6) / 1e9
            print(f"{name} {fmt} B={B}: {ms:8.2f} ms / {steps} steps = {us_step:7.2f} us/step, "
                  f"{B * cfg.n_out * steps / ms * 1e3:9.0f} tokens/s, {B * steps / 50 / (ms * 1e-3):8.0f} audio-s/s, "
                  f"weights {m.weight_bytes_per_step() / 1e6:.1f} MB/step -> {gbs:7.0f} GB/s")
        m.close()


if __name__ == "__main__":
    main()
