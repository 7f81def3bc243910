"""80-dim log-mel fbank + CMN + chunking: float64 CPU restatement.
TEST INFRASTRUCTURE ONLY.

Follows DataGeneratorTrain._logfbank_extractor (data/sr_dataset.py:279-296),
stft/_enframe (simulation/freq_analysis.py:41-150), cmn (reader/preprocess.py:34-41),
_utt2seg (data/sr_dataset.py:40-52), label trim (data/sr_dataset.py:358-363),
GlobalMeanVarianceNormalization.apply_on_ndarray (reader/preprocess.py:211-229).
Pinned against the reference's own code: tests/golden/fbank_*.npz were produced
by oracle/make_golden.py importing /root/reference (tests/test_oracle_fbank.py).
"""
import numpy as np


def mel_matrix(window):
    """window: float32 [80,257] (data/mel80_window.txt) -> mel float32 [257,80].
    data/sr_dataset.py:283-286: columns normalised by their sum, zero sums -> -1."""
    window = np.asarray(window, np.float32)
    t1 = np.sum(window, 0)
    t1[t1 == 0] = -1
    inv = np.diag(1 / t1)
    return window.dot(inv).T


def num_frames(n_samples):
    """T for a wav of n_samples (pre-emphasis drops one sample; last partial frame padded)."""
    m = n_samples - 1
    return max(0, int(-(-(m - 400) // 160)) + 1)


def logfbank(wav, window, dtype=np.float64):
    """wav float32 [n] -> log fbank [T,80] (float64 arithmetic)."""
    mel = mel_matrix(window).astype(dtype)
    wav = np.asarray(wav, np.float32)
    y = (wav[1:] - np.float32(0.96) * wav[:-1]).astype(dtype)
    m = y.shape[0]
    T = num_frames(wav.shape[0])
    need = (T - 1) * 160 + 400
    if need > m:
        y = np.concatenate([y, np.zeros(need - m, dtype)])
    idx = np.arange(T)[:, None] * 160 + np.arange(400)[None, :]
    frames = y[idx] * np.hamming(400)[None, :]
    S = np.fft.rfft(frames, n=512, axis=1)
    P = S.real ** 2 + S.imag ** 2
    fb = P.dot(mel * dtype(32768.0 ** 2)) + 1.0
    return np.log(fb)


def cmn(feat):
    return feat - feat.mean(axis=0, keepdims=True)


def global_mvn(feat, mean_vec, std_vec):
    return (feat - mean_vec) / std_vec


def utt2seg_index(n_fr, seg_len=80, seg_shift=80):
    """Start frames of the chunks _utt2seg produces (tail dropped)."""
    n_seg = int(np.floor((n_fr - seg_len) / seg_shift)) + 1
    return [i * seg_shift for i in range(max(n_seg, 0))]


def chain_subsample_index(T, shift, factor=3):
    """Row indices picked by roll(x, -shift, 1).unfold(1,1,factor) (bin/train_chain.py:251-255)."""
    return [((i * factor) + shift) % T for i in range((T - 1) // factor + 1)]
