"""GPU feature pipeline: waveforms -> log-mel fbank -> CMN/MVN -> chunks or padded batch.

Replaces, for waveform batches already on (or copied to) the device, the CPU NumPy path
DataGeneratorTrain._logfbank_extractor (reference data/sr_dataset.py:279-296) +
stft (simulation/freq_analysis.py:113-150) + cmn (reader/preprocess.py:34-41) +
_utt2seg (data/sr_dataset.py:40-52) + the collate functions' stacking / zero padding
(data/dataloader.py:55-63,94-103) + the chain trainer's roll-and-subsample
(bin/train_chain.py:251-255).  Index maps (which output row reads which frame) are host
numpy, bit-exact; the arithmetic runs in libpk2.so.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from . import mel as _mel

FRAME_LEN, FRAME_SHIFT, FEAT_DIM = 400, 160, 80


def num_frames(n_samples):
    """Frames the reference's stft produces for a wav of n_samples (pre-emphasis drops one
    sample, the last partial frame is zero padded; simulation/freq_analysis.py:64-69)."""
    m = n_samples - 1
    return max(0, -(-(m - FRAME_LEN) // FRAME_SHIFT) + 1)


class FbankExtractor(object):
    """Holds the device-side plan (Hamming window, twiddles, compacted mel matrix)."""

    def __init__(self):
        mel = np.ascontiguousarray(_mel.mel_matrix().astype(np.float32))   # [257, 80]
        h = _lib.vp()
        _lib.check(_lib.lib().pk2_fbank_plan_create(mel.ctypes.data_as(_lib.vp), C.byref(h)),
                   "pk2_fbank_plan_create")
        self._plan = h

    def __del__(self):
        try:
            _lib.lib().pk2_fbank_plan_destroy(self._plan)
        except Exception:
            pass

    def pack(self, wavs):
        """Host side: concatenate float32 waveforms into one pinned buffer + offsets."""
        lens = np.array([len(w) for w in wavs], np.int64)
        woff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        frames = np.array([num_frames(int(n)) for n in lens], np.int64)
        foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)
        buf = torch.empty(int(woff[-1]), dtype=torch.float32).pin_memory()
        bn = buf.numpy()
        for i, w in enumerate(wavs):
            bn[woff[i]:woff[i + 1]] = w
        return buf, woff, foff

    def extract(self, wav_dev, woff, foff):
        """wav_dev: cuda float32 concatenated waveforms.  Returns raw log-fbank [sum T, 80]."""
        _lib.require_cuda(wav_dev, "wav")
        dev = wav_dev.device
        woff_d = torch.from_numpy(np.asarray(woff, np.int64)).to(dev, non_blocking=True)
        foff_d = torch.from_numpy(np.asarray(foff, np.int32)).to(dev, non_blocking=True)
        total = int(foff[-1])
        out = torch.empty(total, FEAT_DIM, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().pk2_fbank(self._plan, _lib.ptr(wav_dev), _lib.ptr(woff_d), _lib.ptr(foff_d),
                                        len(foff) - 1, total, _lib.ptr(out), _lib.stream()), "pk2_fbank")
        return out, foff_d

    def __call__(self, wavs):
        buf, woff, foff = self.pack(wavs)
        wav_dev = buf.cuda(non_blocking=True)
        feats, foff_d = self.extract(wav_dev, woff, foff)
        return feats, foff, foff_d


def utterance_means(feats, foff_d, n_utts):
    mean = torch.empty(n_utts, feats.shape[1], dtype=torch.float32, device=feats.device)
    _lib.check(_lib.lib().pk2_colmean(_lib.ptr(feats), _lib.ptr(foff_d), n_utts, feats.shape[1],
                                      _lib.ptr(mean), _lib.stream()), "pk2_colmean")
    return mean


def gather_norm(feats, row_src, row_utt, mean=None, mvn=None):
    """out[r] = 0 if row_src[r] < 0 else normalised feats[row_src[r]].  row_* are numpy int32."""
    dev = feats.device
    n_rows = len(row_src)
    rs = torch.from_numpy(np.ascontiguousarray(row_src, np.int32)).to(dev, non_blocking=True)
    ru = torch.from_numpy(np.ascontiguousarray(row_utt, np.int32)).to(dev, non_blocking=True)
    out = torch.empty(n_rows, feats.shape[1], dtype=torch.float32, device=dev)
    mm, mi = (None, None) if mvn is None else mvn
    _lib.check(_lib.lib().pk2_gather_norm(_lib.ptr(feats), _lib.ptr(rs), _lib.ptr(ru), _lib.ptr(mean),
                                          _lib.ptr(mm), _lib.ptr(mi), n_rows, feats.shape[1],
                                          _lib.ptr(out), _lib.stream()), "pk2_gather_norm")
    return out


# ------------------------------------------------------------------ index maps (host) ----
def chunk_rows(foff, n_frames=None, seg_len=80, seg_shift=80):
    """Row map of _utt2seg chunking (data/sr_dataset.py:40-52): returns (row_src, row_utt,
    chunk_utt, chunk_start).  n_frames[u] optionally trims each utterance (label trim,
    data/sr_dataset.py:358-363)."""
    src, utt, cu, cs = [], [], [], []
    for u in range(len(foff) - 1):
        T = int(foff[u + 1] - foff[u]) if n_frames is None else int(n_frames[u])
        n_seg = int(np.floor((T - seg_len) / seg_shift)) + 1
        for i in range(max(n_seg, 0)):
            st = i * seg_shift
            src.append(np.arange(st, st + seg_len) + foff[u])
            utt.append(np.full(seg_len, u))
            cu.append(u); cs.append(st)
    if not src:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32)
    return (np.concatenate(src).astype(np.int32), np.concatenate(utt).astype(np.int32),
            np.asarray(cu, np.int32), np.asarray(cs, np.int32))


def padded_rows(foff, n_frames=None, factor=1, shift=0):
    """Row map of SeqDataloader padding (data/dataloader.py:96-103) optionally followed by
    th.roll(x, -shift, 1) + unfold(1, 1, factor) (bin/train_chain.py:251-255):
    out[b, i] = x_padded[b, (i*factor + shift) mod Tmax].  Returns (row_src, row_utt, Tout, num_out)."""
    B = len(foff) - 1
    lens = np.array([int(foff[u + 1] - foff[u]) if n_frames is None else int(n_frames[u]) for u in range(B)])
    Tmax = int(lens.max())
    Tout = (Tmax - 1) // factor + 1
    pos = (np.arange(Tout) * factor + shift) % Tmax                     # frame in the padded batch
    src = np.where(pos[None, :] < lens[:, None], pos[None, :] + np.asarray(foff[:-1])[:, None], -1)
    utt = np.repeat(np.arange(B)[:, None], Tout, axis=1)
    return src.reshape(-1).astype(np.int32), utt.reshape(-1).astype(np.int32), Tout, lens
