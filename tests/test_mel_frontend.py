"""Prompt mel front-end: filterbank restatements against torchaudio's independent implementation of the
librosa/Slaney filterbank (CPU), and the CUDA kernel against the oracle (torch.stft, the reference's own op) on the GPU."""
import numpy as np
import pytest
import torch

import covomix_b200  # noqa: F401
from covomix_b200 import frontend
from oracle import mel_oracle


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", [(8000, 480, 80, 0, 4000), (22050, 1024, 80, 0, 8000), (16000, 400, 40, 50, 7600)])
def test_filterbanks_agree_with_torchaudio(sr, n_fft, n_mels, fmin, fmax):
    torchaudio = pytest.importorskip("torchaudio")
    ref = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax), n_mels, sr, norm="slaney",
                                                mel_scale="slaney").T.numpy()
    a = mel_oracle.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    b = frontend.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    assert a.shape == b.shape == ref.shape == (n_mels, n_fft // 2 + 1)
    assert np.abs(a - ref).max() < 1e-6 * max(1.0, np.abs(ref).max() * 1e2)
    assert np.abs(b - ref).max() < 1e-6 * max(1.0, np.abs(ref).max() * 1e2)


def test_oracle_frame_count_and_known_tone():
    """A pure tone at bin 60 (1 kHz at 8 kHz / 480) puts the mel energy where the filterbank says it should be."""
    n = torch.arange(16000, dtype=torch.float32)
    y = 0.5 * torch.sin(2 * torch.pi * 1000.0 * n / 8000.0)[None]
    mel = mel_oracle.mel_spectrogram(y)
    assert mel.shape == (1, 80, 100)                                   # L / hop frames (generate_mel.py pad = 160)
    fb = mel_oracle.slaney_mel_filterbank(8000, 480, 80, 0, 4000)
    assert int(mel[0, :, 50].argmax()) == int(fb[:, 60].argmax())
    # a bin-centred 0.5-amplitude tone under a periodic hann window: |X[60]| = 0.5 * sum(w) / 2 = 60 and the two
    # neighbouring bins get exactly half of that (the window's own spectrum), everything else is zero
    m = int(fb[:, 60].argmax())
    expect = np.log(fb[m, 59] * 30.0 + fb[m, 60] * 60.0 + fb[m, 61] * 30.0)
    assert abs(float(mel[0, m, 50]) - expect) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(1, 64000), (2, 16000), (3, 8011), (1, 161), (1, 480)])
def test_cuda_mel_matches_oracle(B, L):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(L)
    t = torch.arange(L, dtype=torch.float32) / 8000.0
    y = (0.3 * torch.sin(2 * torch.pi * 440.0 * t)[None] + 0.1 * torch.randn(B, L, generator=g)).clamp(-1, 1)
    ref = mel_oracle.mel_spectrogram(y)
    out = frontend.B200MelSpectrogram(dev)(y.to(dev), 480, 80, 8000, 160, 480, 0, 4000).cpu()
    assert out.shape == ref.shape
    # fp32 direct DFT vs torch's fp32 FFT: log-mel values agree to ~1e-5 absolute; tolerance 1e-4
    assert float((out - ref).abs().max()) < 1e-4


@pytest.mark.gpu
def test_cuda_mel_other_config_and_errors():
    dev = torch.device("cuda:0")
    y = torch.randn(1, 22050, generator=torch.Generator().manual_seed(1)).clamp(-1, 1) * 0.2
    fe = frontend.B200MelSpectrogram(dev)
    out = fe(y.to(dev), 1024, 80, 22050, 256, 1024, 0, 8000).cpu()      # the reference's 22 kHz data-prep setting
    ref = mel_oracle.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
    assert float((out - ref).abs().max()) < 2e-4
    out2 = fe.extract_mel(y[0].to(dev) * 10)                            # clipped like extract_mel (monologue_generation.py:68)
    ref2 = mel_oracle.mel_spectrogram((y * 10).clamp(-1, 1))[0]
    assert float((out2.cpu() - ref2).abs().max()) < 1e-4
    with pytest.raises(ValueError):
        fe(torch.zeros(1, 100, device=dev))
    with pytest.raises(NotImplementedError):
        fe(y.to(dev), center=True)
