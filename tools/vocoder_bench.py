"""HiFi-GAN timing on a B200 at a few (B, T): python tools/vocoder_bench.py   (COVO_HIFIGAN_NO_FUSED=1 for the layered path)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covomix_b200  # noqa: E402,F401
from covomix_b200 import synthetic as syn  # noqa: E402
from covomix_b200.vocoder import B200Generator  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    gen = B200Generator(syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234), syn.HIFIGAN_COVOMIX, dev)
    g = torch.Generator().manual_seed(5)
    shapes = ((8, 1500),) if len(sys.argv) > 1 and sys.argv[1] == "one" else ((1, 256), (8, 1500), (8, 4096), (1, 16384))
    for B, T in shapes:
        mel = syn.synthetic_logmel(g, B, 80, T).to(dev)
        for _ in range(3):
            gen(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gen(mel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"B={B} T={T}: {ms:8.3f} ms  {B * T / 50 / (ms * 1e-3):9.0f} audio-s/s  {281.3e6 * T * B / (ms * 1e-3) / 1e12:6.1f} TFLOP/s "
              f"launches {gen.launches_per_forward()}")


if __name__ == "__main__":
    main()
