"""CPU oracle for the prompt mel front-end.  TEST INFRASTRUCTURE ONLY.

Restates ``mel_spectrogram`` (covomix/util/generate_mel.py:49-72 == data_preparation/generate_mel.py, the function
``extract_mel`` calls, monologue_generation.py:62-74) with the script's constants (monologue_generation.py:349-357:
8 kHz, n_fft = win = 480, hop 160, 80 mels, 0-4000 Hz):
    reflect-pad (n_fft - hop)/2 -> torch.stft(hann(win), center=False, onesided) -> sqrt(re^2 + im^2 + 1e-9)
    -> librosa mel filterbank -> log(clamp(., 1e-5)).

Parity status: the STFT / magnitude / log part calls the very torch ops the reference calls (pinned by construction).
The filterbank is ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` -- third party, absent from this image (librosa is
an unpinned requirement of the reference).  ``slaney_mel_filterbank`` restates its published algorithm (Slaney's Auditory
Toolbox scale: linear below 1 kHz, log above; triangular filters; ``norm='slaney'`` area normalisation) and is checked in
tests/test_mel_frontend.py against torchaudio's independent implementation of the same filterbank
(``melscale_fbanks(norm='slaney', mel_scale='slaney')``) -- "parity unpinned" with respect to librosa itself.
"""
from __future__ import annotations

import numpy as np
import torch


def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` defaults (htk=False, norm='slaney') -> [n_mels, 1 + n_fft//2] f32."""
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


def mel_spectrogram(y: torch.Tensor, n_fft=480, num_mels=80, sampling_rate=8000, hop_size=160, win_size=480, fmin=0,
                    fmax=4000, center=False) -> torch.Tensor:
    """generate_mel.py:49-72.  y [B, L] in [-1, 1] -> [B, num_mels, frames]."""
    basis = torch.from_numpy(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
    window = torch.hann_window(win_size)
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window, center=center, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    spec = torch.view_as_real(spec)
    spec = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)
    spec = torch.matmul(basis, spec)
    return torch.log(torch.clamp(spec, min=1e-5))
