"""The 80x257 mel window the reference loads from data/mel80_window.txt
(data/sr_dataset.py:269-277), regenerated instead of shipped.

The reference's comment (data/sr_dataset.py:271-273) says the file came from
``librosa.filters.mel(16000, 512, n_mels=80, fmax=7690, htk=True)``; the HTK mel
scale with Slaney area normalisation below reproduces the file bit-exactly in
float32 (tests/test_oracle_fbank.py::test_mel_window_regenerated_bit_exact).
"""
import functools

import numpy as np


def _hz2mel(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)


def _mel2hz(m):
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)


@functools.lru_cache(maxsize=None)
def _window(sr=16000, n_fft=512, n_mels=80, fmin=0.0, fmax=7690.0):
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = _mel2hz(np.linspace(_hz2mel(fmin), _hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def mel80_window():
    """float32 [80,257], equal to the reference's data/mel80_window.txt."""
    return _window().copy()


def mel_matrix():
    """float32 [257,80]: (W diag(1/colsum W))^T with zero column sums -> -1
    (data/sr_dataset.py:283-286)."""
    w = mel80_window()
    t1 = np.sum(w, 0)
    t1[t1 == 0] = -1
    inv = np.diag(1 / t1)
    return w.dot(inv).T
