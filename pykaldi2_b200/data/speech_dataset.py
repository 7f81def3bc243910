"""``SpeechDataset(config)``: the reference's corpus description (data yaml -> config["source_paths"], entries with
``type / wav / label / aux_label``; bin/train_ce.py:84-88, example/librispeech/README.md:20-36) served as RAW
waveforms + labels, because feature extraction runs on the GPU here (pipeline.FeaturePipeline).

Items have the layout of ``SyntheticWaveDataset``: ``(wav float32 [n], [utt_id], pdf labels int64 [T, 1],
[transition ids int64 [1, T]])`` and are collated by ``WaveDataloader``.  Differences from the reference's
``data/sr_dataset.py`` (documented, deliberate): the on-the-fly acoustic simulation (noise / RIR mixing,
simulation/*.py) is not applied (SURVEY marks it optional).  Epoch semantics follow the reference:
  * sequence mode (``data_config.sequence_mode``: train_se / train_chain, data/sr_dataset.py:104-107,192-194): an
    epoch visits EVERY utterance; ``sweep_size`` plays no role;
  * chunk mode (train_ce, :55-84,128): an epoch is ``sweep_size`` hours of randomly drawn material: a fresh random
    subset of utterances per epoch (``set_epoch``), not a fixed prefix of the sorted list.
"""
import numpy as np
from torch.utils.data import Dataset

from ..reader import zip_io


class SpeechDataset(Dataset):
    def __init__(self, config):
        self.config = config
        dc = config.get("data_config", {})
        self.load_label = dc.get("load_label", True)
        self.sequence_mode = bool(dc.get("sequence_mode", False))
        self.fs = 16000
        self.reader = zip_io.ZipWaveIO("float32", self.fs)
        self.transform = None                     # kept for the reference's attribute protocol (bin/train_se.py:103-106)
        self.stream_idx_for_transform = [0]
        self.items = []                           # (wav address, utt_id)
        self.labels, self.aux = {}, {}
        sources = config.get("source_paths") or []
        if not sources:
            raise ValueError("SpeechDataset: config['source_paths'] is empty (pass the data yaml with -data)")
        for src in sources:
            names = sorted(self.reader.walk(src["wav"]), key=zip_io.utt_id_of)
            ids = [zip_io.utt_id_of(n) for n in names]
            keep = set(ids)
            lab = aux = None
            if self.load_label and src.get("label"):
                lab = zip_io.read_labels(src["label"], keep)
                keep &= set(lab)
            if self.load_label and src.get("aux_label"):
                aux = zip_io.read_labels(src["aux_label"], keep)
                keep &= set(aux)
            for n, u in zip(names, ids):
                if u in keep:
                    self.items.append((n, u))
                    if lab is not None:
                        self.labels[u] = lab[u]
                    if aux is not None:
                        self.aux[u] = aux[u]
        if not self.items:
            raise ValueError("SpeechDataset: no utterance has both a waveform and labels")
        sweep = config.get("sweep_size")
        self.max_items = None
        if sweep and not self.sequence_mode:       # hours per sweep -> utterances per epoch at LibriSpeech's mean 12.3 s
            self.max_items = max(1, int(float(sweep) * 3600.0 / 12.3))
        self.seed = int(config.get("seed", 0))
        self._subset = None
        self.set_epoch(0)
        # the handles opened while listing the archives must not be inherited by forked DataLoader workers (shared
        # file offset): every process reopens its own on first use
        self.reader.close()

    def set_epoch(self, epoch):
        """Chunk mode with a sweep cap: draw this epoch's random subset of utterances."""
        if self.max_items is not None and self.max_items < len(self.items):
            rng = np.random.default_rng(self.seed * 7919 + int(epoch))
            self._subset = np.sort(rng.choice(len(self.items), self.max_items, replace=False))
        else:
            self._subset = None

    def __len__(self):
        return len(self.items) if self._subset is None else len(self._subset)

    def utt_lengths(self):
        """Frames per utterance for length-balanced sharding: the label length where labels exist (label and feature
        lengths agree up to the trim of data/sr_dataset.py:358-363), else unknown (equal weights)."""
        idx = range(len(self.items)) if self._subset is None else self._subset
        return np.array([len(self.labels[self.items[i][1]]) if self.items[i][1] in self.labels else 1 for i in idx])

    def __getitem__(self, i):
        if self._subset is not None:
            i = int(self._subset[i])
        name, utt = self.items[i]
        _, wav = self.reader.read_wav(name)
        if wav.ndim > 1:
            wav = wav[:, 0]
        lab = self.labels.get(utt)
        aux = self.aux.get(utt)
        pdf = None if lab is None else lab.astype(np.int64)[:, None]
        tid = None if aux is None else [aux.astype(np.int64)[None, :]]
        return np.ascontiguousarray(wav, np.float32), [utt], pdf, tid
