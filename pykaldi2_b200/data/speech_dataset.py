"""``SpeechDataset(config)``: the reference's corpus description (data yaml -> config["source_paths"], entries with
``type / wav / label / aux_label``; bin/train_ce.py:84-88, example/librispeech/README.md:20-36) served as RAW
waveforms + labels, because feature extraction runs on the GPU here (pipeline.FeaturePipeline).

Items have the layout of ``SyntheticWaveDataset``: ``(wav float32 [n], [utt_id], pdf labels int64 [T, 1],
[transition ids int64 [1, T]])`` and are collated by ``WaveDataloader``.  Differences from the reference's
``data/sr_dataset.py`` (documented, deliberate): one epoch visits every labelled utterance once in shuffled order
instead of drawing ``sweep_size`` hours by random sampling (``sweep_size`` caps the epoch length when given), and
the on-the-fly acoustic simulation (noise / RIR mixing, simulation/*.py) is not applied (SURVEY marks it optional).
"""
import numpy as np
from torch.utils.data import Dataset

from ..reader import zip_io


class SpeechDataset(Dataset):
    def __init__(self, config):
        self.config = config
        dc = config.get("data_config", {})
        self.load_label = dc.get("load_label", True)
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
        if sweep:                                  # hours per sweep -> utterance cap at LibriSpeech's mean 12.3 s
            self.max_items = max(1, int(float(sweep) * 3600.0 / 12.3))

    def __len__(self):
        return len(self.items) if self.max_items is None else min(len(self.items), self.max_items)

    def __getitem__(self, i):
        name, utt = self.items[i]
        _, wav = self.reader.read_wav(name)
        if wav.ndim > 1:
            wav = wav[:, 0]
        lab = self.labels.get(utt)
        aux = self.aux.get(utt)
        pdf = None if lab is None else lab.astype(np.int64)[:, None]
        tid = None if aux is None else [aux.astype(np.int64)[None, :]]
        return np.ascontiguousarray(wav, np.float32), [utt], pdf, tid
