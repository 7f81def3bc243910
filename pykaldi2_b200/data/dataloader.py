"""Batch layout contract of the reference's dataloaders (data/dataloader.py:34-136) plus
synthetic datasets for the B200 trainers.

``chunk_collate`` / ``seq_collate`` reproduce ChunkDataloader.collate_fn (:55-63) and
SeqDataloader.collate_fn (:94-136): features zero padded, labels padded with -100, batch dict keys
``utt_ids, x, y[, num_frs, aux]``.  ``ChunkDataloader`` / ``SeqDataloader`` keep the reference's
constructor signatures; the distributed branch shards with torch.distributed's rank/size instead of
Horovod's (data/dataloader.py:45-46,83-84).

``SyntheticWaveDataset`` yields seeded LibriSpeech-shaped waveforms + labels (corpora in the reference's zip /
label-file formats: data/speech_dataset.py); items are raw waveforms because the fbank runs on the GPU
(pipeline.FeaturePipeline), not in the DataLoader workers.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Sampler
from torch.utils.data.distributed import DistributedSampler

from .. import dist as pkdist
from .. import synth
from . import fbank as fb


def chunk_collate(batch):
    feats, utt_ids, labels = zip(*batch)
    return {"utt_ids": utt_ids,
            "x": torch.FloatTensor(np.stack(feats)),
            "y": torch.LongTensor(np.stack(labels))}


def _pad_stack(mats, fill, dtype):
    n_frs = [m.shape[0] for m in mats]
    out = np.full((len(mats), max(n_frs), mats[0].shape[1]), fill, dtype=dtype)
    for i, m in enumerate(mats):
        out[i, :m.shape[0], :] = m
    return out, n_frs


def seq_collate(batch, test_only=False):
    if test_only:
        feats, utt_ids = zip(*batch)
        x, num_frs = _pad_stack(feats, 0, np.float32)
        return {"utt_ids": utt_ids, "num_frs": num_frs, "x": torch.from_numpy(x)}
    feats, utt_ids, labels, aux = zip(*batch)
    y, num_labs = _pad_stack(labels, -100, np.int64)
    x, num_frs = _pad_stack(feats, 0, np.float32)
    assert num_labs == num_frs, "The numbers of frames and labels are not equal"
    return {"utt_ids": utt_ids, "num_frs": num_frs, "x": torch.from_numpy(x), "y": torch.from_numpy(y), "aux": aux}


def _sampler(dataset, distributed):
    if distributed and pkdist.size() > 1:
        return DistributedSampler(dataset, num_replicas=pkdist.size(), rank=pkdist.rank())
    return None


class _EpochMixin(object):
    def set_epoch(self, epoch):
        """Reseed the shuffle of the distributed / balanced sampler (the reference never calls
        DistributedSampler.set_epoch, so every epoch sees the same order: data/dataloader.py:45-53)."""
        for s in (getattr(self, "sampler", None), getattr(self, "batch_sampler", None)):
            if s is not None and hasattr(s, "set_epoch"):
                s.set_epoch(epoch)


class BalancedBatchSampler(Sampler):
    """Batch sampler of the variable-length trainers under data parallelism: an epoch is a seeded permutation of the
    utterances cut into GLOBAL minibatches of batch_size * world utterances; each global minibatch is split across
    the ranks by ``dist.balanced_shards`` (cost = tmax_weight * longest + sum of frames), and a rank yields its own
    share.  All ranks draw the same permutation (seed + epoch), so no communication is needed.  The tail is padded
    by wrap-around like DistributedSampler (drop_last=False)."""

    def __init__(self, lengths, batch_size, world=None, rank=None, seed=0, shuffle=True, tmax_weight=56.0):
        self.lengths = np.asarray(lengths, np.float64)
        self.batch_size = int(batch_size)
        self.world = pkdist.size() if world is None else int(world)
        self.rank = pkdist.rank() if rank is None else int(rank)
        self.seed, self.shuffle, self.tmax_weight, self.epoch = int(seed), bool(shuffle), float(tmax_weight), 0
        self.global_batch = self.batch_size * self.world

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def __len__(self):
        return (len(self.lengths) + self.global_batch - 1) // self.global_batch

    def __iter__(self):
        n = len(self.lengths)
        perm = np.random.default_rng(self.seed + self.epoch).permutation(n) if self.shuffle else np.arange(n)
        total = len(self) * self.global_batch
        if total > n:
            perm = np.concatenate([perm, np.resize(perm, total - n)])
        for g in range(len(self)):
            idx = perm[g * self.global_batch:(g + 1) * self.global_batch]
            shards = pkdist.balanced_shards(self.lengths[idx], self.world, self.batch_size, self.tmax_weight)
            yield [int(idx[j]) for j in shards[self.rank]]


class ChunkDataloader(_EpochMixin, DataLoader):
    def __init__(self, dataset, batch_size, distributed=False, num_workers=0, timeout=1000):
        sampler = _sampler(dataset, distributed)
        super().__init__(dataset, batch_size=batch_size, shuffle=(sampler is None), sampler=sampler,
                         num_workers=num_workers, collate_fn=chunk_collate, drop_last=False,
                         timeout=timeout if num_workers > 0 else 0)


class SeqDataloader(_EpochMixin, DataLoader):
    def __init__(self, dataset, batch_size, num_workers=0, distributed=False, test_only=False, timeout=1000):
        self.test_only = test_only
        sampler = _sampler(dataset, distributed)
        super().__init__(dataset, batch_size=batch_size, shuffle=(sampler is None), sampler=sampler,
                         num_workers=num_workers, collate_fn=lambda b: seq_collate(b, self.test_only),
                         drop_last=False, timeout=timeout if num_workers > 0 else 0)


def wave_collate(batch):
    """Collate for SyntheticWaveDataset items: keep waveforms ragged, labels per utterance."""
    wavs, utt_ids, labels, aux = zip(*batch)
    return {"utt_ids": utt_ids, "wav": list(wavs), "label": list(labels), "aux": list(aux)}


class SyntheticWaveDataset(Dataset):
    """n_utts seeded utterances: (waveform float32, [utt_id], pdf labels int [T,1], [tid labels [1,T]])."""

    def __init__(self, n_utts, num_pdfs, seed=1234, min_dur=1.5, max_dur=30.0):
        self.n, self.num_pdfs, self.seed = int(n_utts), int(num_pdfs), int(seed)
        rng = np.random.default_rng(seed)
        self.durs = synth.make_durations(self.n, rng, min_dur, max_dur)

    def __len__(self):
        return self.n

    def utt_lengths(self):
        """Frames per utterance (for length-balanced sharding), without touching the audio."""
        return np.array([fb.num_frames(int(round(d * 16000))) for d in self.durs])

    def __getitem__(self, i):
        rng = np.random.default_rng(self.seed * 100003 + i)
        wav = synth.make_waveforms([self.durs[i]], rng)[0]
        T = fb.num_frames(len(wav))
        pdf = rng.integers(0, self.num_pdfs, size=(T, 1))
        tid = (2 * pdf[:, 0] + 1 + rng.integers(0, 2, size=T))[None, :]
        return wav, ["synth-%06d" % i], pdf, [tid]


class WaveDataloader(_EpochMixin, DataLoader):
    """``balanced=True`` (the sequence trainers): under data parallelism the global minibatch is split across ranks
    by modelled step time (BalancedBatchSampler) instead of at random; needs ``dataset.utt_lengths()``.
    ``batch_transform``: callable applied to the collated batch dict INSIDE the loader (i.e. in the worker processes
    when num_workers > 0): host-side per-utterance index work such as building the numerator graphs of LF-MMI from
    the alignments belongs there, next to the reference's own CPU data path, not in the training loop."""

    def __init__(self, dataset, batch_size, num_workers=0, distributed=False, timeout=1000, balanced=False, seed=0,
                 batch_transform=None):
        to = timeout if num_workers > 0 else 0
        collate = wave_collate if batch_transform is None else (lambda b: batch_transform(wave_collate(b)))
        if balanced and distributed and pkdist.size() > 1 and hasattr(dataset, "utt_lengths"):
            bs = BalancedBatchSampler(dataset.utt_lengths(), batch_size, seed=seed)
            super().__init__(dataset, batch_sampler=bs, num_workers=num_workers, collate_fn=collate, timeout=to)
            return
        sampler = _sampler(dataset, distributed)
        super().__init__(dataset, batch_size=batch_size, shuffle=(sampler is None), sampler=sampler,
                         num_workers=num_workers, collate_fn=collate, drop_last=False, timeout=to)
