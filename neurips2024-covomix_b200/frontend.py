"""Host-side mirror of the prompt mel front-end: ``mel_spectrogram`` (covomix/util/generate_mel.py:49-72) and the tensor
part of ``extract_mel`` (monologue_generation.py:62-74), backed by libcovomix_b200.so (``covo_mel_forward``: reflect pad,
windowed DFT magnitude, mel filterbank and log in one kernel).  Reading / resampling the prompt file
(``librosa.load``) stays with the caller -- it is I/O, not arithmetic on this path.

The reference takes its filterbank from ``librosa.filters.mel`` (third party, absent here); ``slaney_mel_filterbank``
restates that published construction (Slaney scale, slaney area norm) and is cross-checked against torchaudio's
implementation in tests/test_mel_frontend.py.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import _native as nat


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f * 3.0 / 200.0
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * 27.0 / np.log(6.4), lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0)), m * 200.0 / 3.0)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """``librosa.filters.mel`` with its defaults (htk=False, norm='slaney'): triangular filters on the Slaney mel scale
    (linear to 1 kHz = mel 15, then 27 steps per factor 6.4), each scaled by 2 / bandwidth.  -> f32 [n_mels, n_fft//2 + 1]."""
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    up = (freqs[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
    down = (edges[2:, None] - freqs[None, :]) / (edges[2:] - edges[1:-1])[:, None]
    tri = np.clip(np.minimum(up, down), 0.0, None)
    return (tri * (2.0 / (edges[2:] - edges[:-2]))[:, None]).astype(np.float32)


class B200MelSpectrogram:
    """``mel = B200MelSpectrogram(device)(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax)`` with the
    reference's argument order; handles are cached per parameter set like the reference's ``mel_basis`` / ``hann_window``
    dicts (generate_mel.py:46-58)."""

    def __init__(self, device="cuda:0"):
        self.device = nat.resolve_device(device)
        self._handles: Dict[Tuple, C.c_void_p] = {}

    def _handle(self, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax):
        key = (n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax)
        h = self._handles.get(key)
        if h is None:
            basis = np.ascontiguousarray(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
            window = torch.hann_window(win_size).numpy().astype(np.float32)
            cfg = nat.MelCfg(n_fft=n_fft, hop_size=hop_size, win_size=win_size, num_mels=num_mels)
            h = C.c_void_p()
            nat.check(nat.lib().covo_mel_create(C.byref(cfg), window.ctypes.data_as(C.c_void_p),
                                                basis.ctypes.data_as(C.c_void_p), self.device.index, C.byref(h)),
                      "covo_mel_create")
            self._handles[key] = h
        return h

    @torch.inference_mode()
    def __call__(self, y, n_fft=480, num_mels=80, sampling_rate=8000, hop_size=160, win_size=480, fmin=0, fmax=4000,
                 center=False):
        if center:
            raise NotImplementedError("covomix_b200: the reference calls mel_spectrogram with center=False only")
        y = y.to(device=self.device, dtype=torch.float32).contiguous()
        if y.ndim != 2:
            raise ValueError(f"y must be [B, L], got {tuple(y.shape)}")
        B, L = y.shape
        h = self._handle(int(n_fft), int(num_mels), int(sampling_rate), int(hop_size), int(win_size), float(fmin), float(fmax))
        T = nat.lib().covo_mel_frames(h, L)
        if T < 1:
            raise ValueError(f"signal of {L} samples is too short for n_fft={n_fft}, hop={hop_size}")
        mel = torch.empty(B, num_mels, T, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib().covo_mel_forward(h, C.c_void_p(y.data_ptr()), C.c_void_p(mel.data_ptr()), B, L,
                                             C.c_void_p(stream)), "covo_mel_forward")
        return mel

    def extract_mel(self, wav: torch.Tensor) -> torch.Tensor:
        """Tensor part of ``extract_mel`` (monologue_generation.py:68-74): clip to [-1, 1], mel of the 1-D signal -> [80, T]."""
        return self(wav.reshape(1, -1).clamp(-1, 1))[0]

    def close(self):
        for h in self._handles.values():
            nat.lib().covo_mel_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
