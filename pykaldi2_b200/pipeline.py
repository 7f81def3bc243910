"""One training step of the hot path, composed from the reference-facing pieces
(feature pipeline -> LSTMAM -> ops.ops loss -> NCCL gradient averaging -> clip -> optimizer).

Used by bench.py and by the trainers in bin/; it is the body of the reference's
run_train_epoch loops (bin/train_ce.py:177-208, bin/train_se.py:222-277,
bin/train_chain.py:244-308) with the CPU data path and the per-utterance host round trips
replaced by the kernels of libpk2.so.
"""
import numpy as np
import torch
import torch.nn as nn

from .data import fbank as fb
from .ops import ops


class FeaturePipeline(object):
    """waveforms (pinned host or device) -> normalised, padded / chunked / subsampled batch."""

    def __init__(self, use_cmn=True, mvn=None):
        self.ex = fb.FbankExtractor()
        self.use_cmn = use_cmn
        self.mvn = mvn

    def features(self, wav, woff, foff):
        wav_dev = wav if wav.is_cuda else wav.cuda(non_blocking=True)
        feats, foff_d = self.ex.extract(wav_dev, woff, foff)
        mean = fb.utterance_means(feats, foff_d, len(foff) - 1) if self.use_cmn else None
        return feats, mean

    def sequence_batch(self, wav, woff, foff, factor=1, shift=0, n_frames=None):
        """-> (x [B, Tout, 80], input lengths per utterance)"""
        feats, mean = self.features(wav, woff, foff)
        src, utt, Tout, lens = fb.padded_rows(foff, n_frames, factor, shift)
        x = fb.gather_norm(feats, src, utt, mean, self.mvn).view(len(foff) - 1, Tout, fb.FEAT_DIM)
        return x, lens

    def chunk_batch(self, wav, woff, foff, seg_len=80, seg_shift=80, n_frames=None):
        """-> (x [n_chunks, seg_len, 80], chunk_utt, chunk_start)"""
        feats, mean = self.features(wav, woff, foff)
        src, utt, cu, cs = fb.chunk_rows(foff, n_frames, seg_len, seg_shift)
        x = fb.gather_norm(feats, src, utt, mean, self.mvn).view(-1, seg_len, fb.FEAT_DIM)
        return x, cu, cs


class ChunkPool(object):
    """Device-side stand-in for the reference's DataBuffer (data/sr_dataset.py:55-84: a 20 000-sample buffer that the
    chunk generator keeps filled and from which training samples are popped AT RANDOM).  Chunks cut from whole
    utterances on the GPU are appended; ``draw(n)`` removes and returns n chunks chosen uniformly at random (host
    RNG, seeded), so a minibatch mixes chunks of many utterances and no chunk is ever discarded."""

    def __init__(self, capacity, seg_len, feat_dim, device, seed=0):
        self.capacity = int(capacity)
        self.x = torch.empty(self.capacity, seg_len, feat_dim, dtype=torch.float32, device=device)
        self.y = torch.empty(self.capacity, seg_len, dtype=torch.int64, device=device)
        self.n = 0
        self.rng = np.random.default_rng(seed)

    def room(self):
        return self.capacity - self.n

    def add(self, x, y):
        k = x.shape[0]
        if k > self.room():
            raise RuntimeError("ChunkPool overflow: %d chunks, room for %d" % (k, self.room()))
        self.x[self.n:self.n + k] = x
        self.y[self.n:self.n + k] = y
        self.n += k

    def draw(self, n):
        n = min(int(n), self.n)
        pick = self.rng.choice(self.n, size=n, replace=False)
        dev = self.x.device
        idx = torch.from_numpy(pick.astype(np.int64)).to(dev)
        bx, by = self.x[idx], self.y[idx]
        # fill the holes below the new end with the surviving entries of the tail
        new_n = self.n - n
        picked = np.zeros(self.n, bool)
        picked[pick] = True
        holes = pick[pick < new_n]
        tail = np.nonzero(~picked[new_n:])[0] + new_n
        if len(holes):
            h = torch.from_numpy(np.sort(holes).astype(np.int64)).to(dev)
            t = torch.from_numpy(tail.astype(np.int64)).to(dev)
            self.x[h] = self.x[t]
            self.y[h] = self.y[t]
        self.n = new_n
        return bx, by


def ce_loss(logits, labels, reduction="mean"):
    """nn.CrossEntropyLoss(ignore_index=-100) on the fused kernel (bin/train_ce.py:134,189)."""
    return _CE.apply(logits, labels, reduction)


class _CE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, reduction):
        from . import _lib
        _lib.require_cuda(logits, "logits")
        N = logits.shape[-1]
        lg = logits.detach().contiguous().view(-1, N)
        lab = labels.contiguous().view(-1).to(torch.int64)
        R = lg.shape[0]
        n_valid = (lab >= 0).sum().clamp(min=1).to(torch.float32)
        rows = torch.empty(R, dtype=torch.float32, device=lg.device)
        grad = torch.empty_like(lg)
        _lib.check(_lib.lib().pk2_ce_softmax(_lib.ptr(lg), _lib.ptr(lab), R, N, 1.0, _lib.ptr(rows), _lib.ptr(grad),
                                             _lib.stream()), "pk2_ce_softmax")
        tot = rows.sum()
        if reduction == "mean":
            ctx.scale = 1.0 / n_valid
            tot = tot / n_valid
        else:
            ctx.scale = None
        ctx.save_for_backward(grad)
        ctx.shape = logits.shape
        return tot

    @staticmethod
    def backward(ctx, grad_out):
        grad, = ctx.saved_tensors
        g = grad * (grad_out if ctx.scale is None else grad_out * ctx.scale)
        return g.view(ctx.shape), None, None


def finish_step(model, optimizer, averager, max_grad_norm):
    """allreduce(mean) -> clip -> optimizer step (bin/train_ce.py:192-196 with Horovod folded in)."""
    if averager is not None:
        averager.average()
    norm = nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return norm


class PendingValue(object):
    """A scalar on its way from the device to pinned host memory.  ``value()`` waits for THAT copy only, so a training
    loop can read the loss of step k after it has enqueued step k+1: the host then runs one step ahead of the GPU and
    its launch work (index maps, ~150 launches) never leaves the GPU idle at the start of a step."""

    def __init__(self, t):
        self._host = torch.empty((), dtype=t.dtype, pin_memory=True)
        self._host.copy_(t.detach(), non_blocking=True)
        self._ev = torch.cuda.Event()
        self._ev.record()
        self.d2h_bytes = self._host.element_size()

    def value(self):
        self._ev.synchronize()
        return float(self._host)


def chain_step(model, optimizer, averager, feat, den_graph, chain_opts, wav, woff, foff, supervisions,
               epoch=0, max_grad_norm=5.0, factor=3, after_backward=None, events=None, sync=True):
    """One LF-MMI step (bin/train_chain.py:244-292).  Returns (objf, total input frames); objf is a float, or with
    ``sync=False`` a PendingValue (read it with .value() after the NEXT step has been enqueued).
    ``after_backward``: host callback run once the backward pass is enqueued (e.g. prefetch of the next batch).
    ``events``: list that receives (start, backward done) CUDA events of the step -- the rank's own compute time,
    before it waits for the other ranks in the gradient all-reduce (bench.py: per-rank spread)."""
    if events is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    shift = epoch % factor                                   # frame_shift = -(epoch % 3) then roll
    x, lens = feat.sequence_batch(wav, woff, foff, factor=factor, shift=shift)
    # the output layer runs on the frames the loss reads (one supervision frame per output frame), not on the padding
    valid = supervisions.num_frames_host if hasattr(supervisions, "num_frames_host") else \
        [s.frames_per_sequence for s in supervisions]
    prediction = model(x, valid_lengths=valid)
    loss = ops.ChainObjtiveFunction.apply_batch(prediction, den_graph, supervisions, chain_opts)
    pending = None if sync else PendingValue(loss)
    loss.backward()
    if events is not None:
        ev1.record()
        events.append((ev0, ev1))
    if after_backward is not None:
        after_backward()
    finish_step(model, optimizer, averager, max_grad_norm)
    return (float(loss.item()) if sync else pending), int(np.sum(lens))
