"""Host-side graph objects of the hot path: the duck-typed stand-ins for the PyKaldi
objects the reference passes through ``ops.ops``.

  DenominatorGraph      <- kaldi_chain.DenominatorGraph(den_fst, num_pdfs)   bin/train_chain.py:167,202
  Supervision           <- kaldi_chain.Supervision (numerator FST)            bin/train_chain.py:271-272
  ChainTrainingOptions  <- kaldi_chain.ChainTrainingOptions                   bin/train_chain.py:191-193
  Lattice               <- the decoder's lattice, decode_out["lattice"]       ops/ops.py:55-56
  TidPdfMap             <- kaldi_hmm.TransitionModel (tid -> pdf)             bin/train_se.py:164-170
  SyntheticLatticeProvider <- MappedLatticeFasterRecognizer (.decode)         bin/train_se.py:179-181

All index tensors (CSR arrays, state times, level offsets, drop masks) are built here with
numpy, bit-exactly reproducible, and uploaded once; the floating-point work is CUDA.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _csr(keys, n):
    """CSR offsets [n + 1] of `keys` (row of every entry); bincount, not np.add.at (10x faster)."""
    cnt = np.bincount(np.asarray(keys, np.int64), minlength=n) if len(keys) else np.zeros(n, np.int64)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=off[1:])
    return off.astype(np.int32)


class ChainTrainingOptions(object):
    """Same attribute names as kaldi_chain.ChainTrainingOptions."""

    def __init__(self, leaky_hmm_coefficient=1e-5, xent_regularize=0.0, l2_regularize=0.0,
                 out_of_range_regularize=0.01):
        self.leaky_hmm_coefficient = leaky_hmm_coefficient
        self.xent_regularize = xent_regularize
        self.l2_regularize = l2_regularize
        self.out_of_range_regularize = out_of_range_regularize


class DenominatorGraph(object):
    """Denominator graph of the chain objective, resident on the current CUDA device.

    ``fst``: dict(num_states, start, src, dst, ilabel (pdf+1), weight (cost), final (cost)),
    arcs sorted by source state, epsilon-free.  Holds Kaldi's arrays
    (forward_transitions / backward_transitions as CSR, initial_probs) as numpy for
    inspection, and the device handle used by the kernels.
    """

    def __init__(self, fst, num_pdfs):
        S = int(fst["num_states"])
        src = np.asarray(fst["src"], np.int64)
        dst = np.asarray(fst["dst"], np.int64)
        pdf = np.asarray(fst["ilabel"], np.int64) - 1
        if len(src) and (np.diff(src) < 0).any():
            raise ValueError("den fst arcs must be sorted by source state")
        if (pdf < 0).any() or (pdf >= num_pdfs).any():
            raise ValueError("den fst must be epsilon-free with ilabel = pdf+1 < num_pdfs+1")
        prob = np.exp(-np.asarray(fst["weight"], np.float64)).astype(np.float32)
        self._S, self._N = S, int(num_pdfs)
        self.fwd_off = _csr(src, S)
        self.fwd_prob = prob
        self.fwd_pdf = pdf.astype(np.int32)
        self.fwd_state = dst.astype(np.int32)
        order = np.argsort(dst, kind="stable")
        self.bwd_off = _csr(dst, S)
        self.bwd_prob = prob[order]
        self.bwd_pdf = pdf[order].astype(np.int32)
        self.bwd_state = src[order].astype(np.int32)
        self.initial_probs = self._initial_probs(fst, src, dst)
        self._handle = _lib.vp(0)
        h = _lib.vp()
        keep = [np.ascontiguousarray(a) for a in
                (self.fwd_off, self.fwd_prob, self.fwd_pdf, self.fwd_state, self.initial_probs)]
        _lib.check(_lib.lib().pk2_den_graph_create(
            S, self._N, *[a.ctypes.data_as(_lib.vp) for a in keep], C.byref(h)), "pk2_den_graph_create")
        self._handle = h

    @staticmethod
    def _initial_probs(fst, src, dst):
        # Kaldi DenominatorGraph::SetInitialProbs: 100 steps from the start state with
        # normalised transition probs, averaged (double), then cast to float.
        S = int(fst["num_states"])
        probd = np.exp(-np.asarray(fst["weight"], np.float64))
        tot = np.exp(-np.asarray(fst["final"], np.float64))
        np.add.at(tot, src, probd)
        step = probd / tot[src]
        cur = np.zeros(S)
        cur[int(fst["start"])] = 1.0
        avg = np.zeros(S)
        for _ in range(100):
            avg += cur / 100.0
            nxt = np.zeros(S)
            np.add.at(nxt, dst, cur[src] * step)
            cur = nxt / nxt.sum()
        return avg.astype(np.float32)

    @classmethod
    def from_file(cls, path, num_pdfs):
        """``den.fst`` as Kaldi writes it (OpenFst binary, vector / standard) or in fstprint text form
        (reference bin/train_chain.py:196-202 reads it through PyKaldi)."""
        from .reader import fst_io
        return cls(fst_io.read_fst(path), num_pdfs)

    def num_states(self):
        return self._S

    def num_pdfs(self):
        return self._N

    @property
    def handle(self):
        return self._handle

    def set_sm_budget(self, max_clusters=0, reserve_sms=0):
        """Limit later forward-backward calls on this graph to ``max_clusters`` resident clusters of 8 CTAs
        (0 = all that fit) and keep ``reserve_sms`` SMs free of single-CTA kernels, for callers that run
        other kernels next to the denominator (pipeline.chain_step_overlapped)."""
        _lib.check(_lib.lib().pk2_den_set_sm_budget(self._handle, int(max_clusters), int(reserve_sms)),
                   "pk2_den_set_sm_budget")

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().pk2_den_graph_destroy(self._handle)
        except Exception:
            pass


def _level_sort(num_states, start, src, dst, advance, times=None):
    """State times + renumbering that orders states by (time, original index), start first.

    advance[k] = 1 for label-consuming arcs, 0 for epsilon arcs.  Returns
    (times_sorted int32 [S], perm (new -> old), inv (old -> new)).
    """
    S = int(num_states)
    if times is None:
        times = np.full(S, -1, np.int64)
        times[start] = 0
        order = np.argsort(src, kind="stable")
        so, do, ao = src[order], dst[order], advance[order]
        off = _csr(so, S)
        frontier = np.array([start])
        # level-synchronous BFS (epsilon arcs go to same-time states with a higher index)
        while len(frontier):
            nxt = []
            for s in frontier:
                for k in range(off[s], off[s + 1]):
                    d, nt = do[k], times[s] + ao[k]
                    if times[d] < 0:
                        times[d] = nt
                        nxt.append(d)
                    elif times[d] != nt:
                        raise ValueError("paths of different length reach state %d" % d)
            frontier = np.array(sorted(set(nxt)), dtype=np.int64)
        if (times < 0).any():
            raise ValueError("graph has unreachable states")
    times = np.asarray(times, np.int64)
    key = times * 2 + (np.arange(S) != start)
    perm = np.argsort(key * (S + 1) + np.arange(S), kind="stable")
    inv = np.empty(S, np.int64)
    inv[perm] = np.arange(S)
    return times[perm].astype(np.int32), perm, inv


class Supervision(object):
    """Numerator graph of one sequence (what proto_supervision_to_supervision returns).

    ``fst``: epsilon-free acyclic FST dict (ilabel = pdf+1), optional ``state_times``.
    Attributes mirror kaldi_chain.Supervision: weight, num_sequences, frames_per_sequence,
    label_dim.
    """

    def __init__(self, fst, frames_per_sequence, label_dim, weight=1.0):
        self.weight = float(weight)
        self.num_sequences = 1
        self.frames_per_sequence = int(frames_per_sequence)
        self.label_dim = int(label_dim)
        S = int(fst["num_states"])
        src = np.asarray(fst["src"], np.int64)
        dst = np.asarray(fst["dst"], np.int64)
        pdf = np.asarray(fst["ilabel"], np.int64) - 1
        if (pdf < 0).any() or (pdf >= label_dim).any():
            raise ValueError("supervision fst must be epsilon-free with pdf < label_dim")
        w = np.asarray(fst["weight"], np.float32)
        times, perm, inv = _level_sort(S, int(fst["start"]), src, dst, np.ones(len(src), np.int64),
                                       fst.get("state_times"))
        T = self.frames_per_sequence
        fin = np.asarray(fst["final"], np.float32)[perm]
        if (times[np.isfinite(fin)] != T).any() or times.max() != T:
            raise ValueError("supervision final states must all have time frames_per_sequence")
        src, dst = inv[src], inv[dst]
        o = np.argsort(src, kind="stable")
        self.num_states = S
        self.state_time = times
        self.final_cost = fin
        self.level_off = _csr(times, T + 1)          # [T+2]
        self.out_off = _csr(src, S)
        self.out_dst = dst[o].astype(np.int32)
        self.out_pdf = pdf[o].astype(np.int32)
        self.out_w = w[o]
        i = np.argsort(dst, kind="stable")
        self.in_off = _csr(dst, S)
        self.in_src = src[i].astype(np.int32)
        self.in_pdf = pdf[i].astype(np.int32)
        self.in_w = w[i]

    @classmethod
    def from_file(cls, path, frames_per_sequence, label_dim, weight=1.0):
        """Numerator FST of one utterance from a file (OpenFst binary vector / standard, or fstprint text):
        epsilon-free, ilabel = pdf + 1, e.g. the ``fst`` member of a Kaldi chain Supervision dumped with fstprint."""
        from .reader import fst_io
        return cls(fst_io.read_fst(path), frames_per_sequence, label_dim, weight)


def _cat_off(offs):
    """Concatenate per-item CSR offset arrays into one (dropping the shared boundaries)."""
    out, base = [np.zeros(1, np.int32)], 0
    for o in offs:
        out.append(o[1:] + base)
        base += int(o[-1])
    return np.concatenate(out).astype(np.int32)


class SupervisionBatch(object):
    """Device-resident concatenation of several Supervisions (one C-ABI call for the batch)."""

    def __init__(self, sups, device=None):
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.n_seq = len(sups)
        self.weights = [s.weight for s in sups]
        ns = np.array([s.num_states for s in sups], np.int64)
        sbase = np.concatenate([[0], np.cumsum(ns)])
        abase = np.concatenate([[0], np.cumsum([len(s.out_dst) for s in sups])])
        self.total_states = int(sbase[-1])
        self.total_arcs = int(sum(len(s.out_dst) for s in sups))
        self.num_frames_host = [s.frames_per_sequence for s in sups]
        lvl_base = np.concatenate([[0], np.cumsum([len(s.level_off) for s in sups])])[:-1]
        host = {
            "seq_state_off": sbase.astype(np.int32),
            "lvl_base": lvl_base.astype(np.int32),
            "level_off": np.concatenate([s.level_off + np.int32(sbase[i]) for i, s in enumerate(sups)]).astype(np.int32),
            "num_frames": np.array(self.num_frames_host, np.int32),
            "out_off": _cat_off([s.out_off for s in sups]),
            "out_dst": np.concatenate([s.out_dst + np.int32(sbase[i]) for i, s in enumerate(sups)]).astype(np.int32),
            "out_pdf": np.concatenate([s.out_pdf for s in sups]),
            "out_w": np.concatenate([s.out_w for s in sups]),
            "in_off": _cat_off([s.in_off for s in sups]),
            "in_src": np.concatenate([s.in_src + np.int32(sbase[i]) for i, s in enumerate(sups)]).astype(np.int32),
            "in_pdf": np.concatenate([s.in_pdf for s in sups]),
            "in_w": np.concatenate([s.in_w for s in sups]),
            "final_cost": np.concatenate([s.final_cost for s in sups]),
            "state_time": np.concatenate([s.state_time for s in sups]),
        }
        del abase
        # one pinned staging buffer + one H2D copy for all index arrays
        self._dev, self._keep = _upload(host, device)
        self.struct = _lib.SupBatch()
        self.struct.n_seq = self.n_seq
        for k, t in self._dev.items():
            setattr(self.struct, k, t.data_ptr())
        self.h2d_bytes = sum(v.nbytes for v in host.values())


def _upload(host, device):
    """Pack the arrays into one pinned buffer, copy once, return 16-byte aligned views."""
    offs, total = {}, 0
    for k, v in host.items():
        offs[k] = total
        total += (v.nbytes + 15) // 16 * 16
    stage = torch.empty(max(total, 16), dtype=torch.uint8).pin_memory()
    sn = stage.numpy()
    for k, v in host.items():
        sn[offs[k]:offs[k] + v.nbytes] = np.ascontiguousarray(v).view(np.uint8).reshape(-1)
    dev = stage.to(device, non_blocking=True)
    tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
           np.dtype(np.uint8): torch.uint8, np.dtype(np.int64): torch.int64}
    views = {k: dev[offs[k]:offs[k] + v.nbytes].view(tdt[v.dtype]) for k, v in host.items()}
    return views, (dev, stage)     # keep `stage` alive until the async copy has run


class TidPdfMap(object):
    """Minimal TransitionModel stand-in: transition-id -> pdf-id, and (for MPFE and the silence classes of
    sMBR, ops/ops.py:130-147) transition-id -> phone."""

    def __init__(self, tid2pdf, tid2phone=None):
        self.tid2pdf = np.asarray(tid2pdf, np.int32)
        self._num_pdfs = int(self.tid2pdf.max()) + 1
        self.tid2phone = None if tid2phone is None else np.asarray(tid2phone, np.int32)
        if self.tid2phone is not None and len(self.tid2phone) != len(self.tid2pdf):
            raise ValueError("tid2phone and tid2pdf must have one entry per transition id (index 0 unused)")

    @classmethod
    def from_kaldi_text(cls, path):
        """From a Kaldi transition model file: text form (``copy-transition-model --binary=false final.mdl -``) or the
        binary ``final.mdl`` itself (its transition-model head; reader/kaldi_io.py).  The reference reads it through
        PyKaldi (bin/train_se.py:164-170)."""
        from .reader import kaldi_io
        tm = kaldi_io.read_transition_model(path)
        return cls(tm["tid2pdf"], tm["tid2phone"])

    def transition_id_to_pdf(self, tid):
        return int(self.tid2pdf[tid])

    def transition_id_to_phone(self, tid):
        if self.tid2phone is None:
            raise RuntimeError("this TidPdfMap was built without a transition-id -> phone table")
        return int(self.tid2phone[tid])

    def num_pdfs(self):
        return self._num_pdfs

    def num_transition_ids(self):
        return len(self.tid2pdf) - 1


class Lattice(object):
    """A decoding lattice prepared for the level-synchronous kernel.

    ``lat``: dict(num_states, src, dst, tid (0 = epsilon), graph_cost, final_cost), start
    state 0, arcs sorted by src, topologically sorted.  ``state_times`` follows Kaldi's
    LatticeStateTimes.  States are renumbered by (time, index); epsilon arcs are kept in a
    per-level list in topological order of their source.
    """

    def __init__(self, lat):
        S = int(lat["num_states"])
        src = np.asarray(lat["src"], np.int64)
        dst = np.asarray(lat["dst"], np.int64)
        tid = np.asarray(lat["tid"], np.int64)
        gc = np.asarray(lat["graph_cost"], np.float32)
        if (src >= dst).any():
            raise ValueError("lattice must be topologically sorted (src < dst on every arc)")
        adv = (tid != 0).astype(np.int64)
        times, perm, inv = _level_sort_lattice(S, src, dst, adv)
        self.state_times_orig = np.empty(S, np.int32)
        self.state_times_orig[perm] = times          # Kaldi's LatticeStateTimes in input numbering
        self.num_frames = int(times.max())
        T = self.num_frames
        fin = np.asarray(lat["final_cost"], np.float32)[perm]
        if (times[np.isfinite(fin)] != T).any():
            raise ValueError("lattice final states must have time T")
        src, dst = inv[src], inv[dst]
        ne = tid != 0
        s1, d1, t1, g1 = src[ne], dst[ne], tid[ne], gc[ne]
        o = np.argsort(s1, kind="stable")
        self.num_states = S
        self.state_time = times
        self.final_cost = fin
        self.level_off = _csr(times, T + 1)
        self.out_off = _csr(s1, S)
        self.out_dst = d1[o].astype(np.int32)
        self.out_tid = t1[o].astype(np.int32)
        self.out_gc = g1[o]
        i = np.argsort(d1, kind="stable")
        # in-arc k is out-arc _in_from_out[k]; frame of every out-arc (host index work reused by frame_acc)
        rank_out = np.empty(len(o), np.int64)
        rank_out[o] = np.arange(len(o))
        self._in_from_out = rank_out[i].astype(np.int32)
        self._out_frame = times[s1[o]].astype(np.int32)
        self.in_off = _csr(d1, S)
        self.in_src = s1[i].astype(np.int32)
        self.in_tid = t1[i].astype(np.int32)
        self.in_gc = g1[i]
        es, ed, eg = src[~ne], dst[~ne], gc[~ne]
        eo = np.argsort(es, kind="stable")           # new numbering keeps topological order
        es, ed, eg = es[eo], ed[eo], eg[eo]
        self.eps_off = _csr(times[es] if len(es) else np.zeros(0, np.int64), T + 1)
        self.eps_src = es.astype(np.int32)
        self.eps_dst = ed.astype(np.int32)
        self.eps_gc = eg.astype(np.float32)
        # tids present on each frame (for the drop_frames test), as sorted unique (t, tid) keys
        self._frame_tid_keys = _sorted_unique(times[s1].astype(np.int64) * (1 << 32) + t1)

    def frame_acc(self, num_ali, tid2pdf, tid2phone, criterion, silence_phones, one_silence_class=True):
        """Per-arc frame accuracy (uint8 0/1) of Kaldi's LatticeForwardBackwardMpeVariants for the non-epsilon
        arcs, in in-arc and out-arc order: smbr compares the pdf of the arc with the pdf of the reference
        alignment at the arc's frame, mpfe the phones; with one_silence_class (what ops/ops.py:138 passes) an
        arc also counts when both phones are silence phones.  Pure index work (bit-exact)."""
        if criterion not in ("smbr", "mpfe"):
            raise ValueError("criterion must be 'smbr' or 'mpfe' (got %r)" % (criterion,))
        num_ali = np.asarray(num_ali, np.int64)
        if len(num_ali) != self.num_frames:
            raise ValueError("alignment length %d != lattice frames %d" % (len(num_ali), self.num_frames))
        # per-transition-id tables (class to compare, silence flag), then ONE pass over the arcs in out-arc order;
        # the in-arc order is a stored permutation of it
        tid2phone = np.asarray(tid2phone, np.int64)
        cls = (tid2phone if criterion == "mpfe" else np.asarray(tid2pdf, np.int64)).astype(np.int32)
        sil_phone = np.zeros(int(tid2phone.max()) + 2, bool)
        for p in silence_phones:
            if 0 <= int(p) < len(sil_phone):
                sil_phone[int(p)] = True
        tid_sil = sil_phone[np.clip(tid2phone, 0, len(sil_phone) - 1)]
        ref = num_ali[self._out_frame]                         # reference transition id at every arc's frame
        tid = self.out_tid
        same = cls[tid] == cls[ref]
        if one_silence_class:
            out = same | (tid_sil[tid] & tid_sil[ref])
        else:
            out = same & ~tid_sil[tid]
        out = out.astype(np.uint8)
        return out[self._in_from_out], out

    def keep_mask(self, num_ali):
        """1 where the alignment's tid occurs among the lattice's tids of that frame
        (frames where numerator and denominator posteriors are disjoint are dropped)."""
        num_ali = np.asarray(num_ali, np.int64)
        keys = np.arange(len(num_ali), dtype=np.int64) * (1 << 32) + num_ali
        pos = np.searchsorted(self._frame_tid_keys, keys)
        pos = np.minimum(pos, len(self._frame_tid_keys) - 1)
        return (self._frame_tid_keys[pos] == keys).astype(np.uint8)


def _sorted_unique(a):
    """np.unique for small integer arrays without its per-call overhead (called once per lattice level)."""
    if len(a) < 2:
        return a
    a = np.sort(a)
    return a[np.concatenate(([True], a[1:] != a[:-1]))]


def _expand(off, nodes):
    """Indices of the CSR rows `nodes` laid end to end."""
    cnt = off[nodes + 1] - off[nodes]
    tot = int(cnt.sum())
    if tot == 0:
        return np.zeros(0, np.int64)
    return np.repeat(off[nodes], cnt) + (np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt))


def _level_sort_lattice(S, src, dst, adv):
    """Vectorised LatticeStateTimes for topologically sorted lattices (arcs sorted by src): a frontier walk, one
    numpy step per frame.  Label-consuming arcs lead to the next frame; epsilon arcs (adv == 0) stay inside the
    frame and are followed to closure before the frame advances."""
    times = np.full(S, -1, np.int64)
    times[0] = 0
    is_adv = adv.astype(bool)
    all_adv = bool(is_adv.all())
    off_a = _csr(src[is_adv], S).astype(np.int64)
    dst_a = dst[is_adv]
    if not all_adv:
        off_e = _csr(src[~is_adv], S).astype(np.int64)
        dst_e = dst[~is_adv]
    frontier = np.array([0], np.int64)
    t = 0
    while len(frontier):
        if not all_adv:                       # epsilon closure of the frame
            new = frontier
            while len(new):
                d = _sorted_unique(dst_e[_expand(off_e, new)])
                if len(d) and ((times[d] >= 0) & (times[d] != t)).any():
                    raise ValueError("paths of different length reach one lattice state")
                new = d[times[d] < 0]
                times[new] = t
                if len(new):
                    frontier = np.concatenate([frontier, new])
        d = _sorted_unique(dst_a[_expand(off_a, frontier)])
        if len(d) and (times[d] >= 0).any():
            raise ValueError("paths of different length reach one lattice state")
        t += 1
        times[d] = t
        frontier = d
    if (times < 0).any():
        raise ValueError("lattice has unreachable states")
    perm = np.argsort(times * (S + 1) + np.arange(S), kind="stable")
    inv = np.empty(S, np.int64)
    inv[perm] = np.arange(S)
    return times[perm].astype(np.int32), perm, inv


class LatticeBatch(object):
    """Device-resident concatenation of lattices + alignments for pk2_latfb_mmi."""

    def __init__(self, lats, tid2pdf, num_alis, device=None, mpe=None):
        """``mpe``: None, or (criterion, tid2phone, silence_phones) to also upload the per-arc frame accuracies
        pk2_latfb_mpe needs (sMBR / MPFE)."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.n_seq = len(lats)
        ns = np.array([l.num_states for l in lats], np.int64)
        sbase = np.concatenate([[0], np.cumsum(ns)])
        self.total_states = int(sbase[-1])
        self.total_arcs = int(sum(len(l.in_src) for l in lats))
        self.num_frames_host = [l.num_frames for l in lats]
        self.total_frames = int(sum(self.num_frames_host))
        # every index the kernels dereference is validated here, once, on the host: transition ids
        # within the tid -> pdf map, pdfs within the prediction's columns (checked against N at call time)
        t2p = np.asarray(tid2pdf, np.int32)
        self.max_pdf = -1
        for l, a in zip(lats, num_alis):
            for name, tids in (("lattice arc", l.out_tid), ("alignment", np.asarray(a))):
                if len(tids) and (tids.min() < 1 or tids.max() >= len(t2p)):
                    raise ValueError("%s transition id outside 1..%d" % (name, len(t2p) - 1))
                if len(tids):
                    pd = t2p[tids]
                    if pd.min() < 0:
                        raise ValueError("%s transition id maps to no pdf" % name)
                    self.max_pdf = max(self.max_pdf, int(pd.max()))
        for l, a in zip(lats, num_alis):
            if len(a) != l.num_frames:
                raise ValueError("alignment length %d != lattice frames %d" % (len(a), l.num_frames))
        lvl_base = np.concatenate([[0], np.cumsum([len(l.level_off) for l in lats])])[:-1]
        ebase = np.concatenate([[0], np.cumsum([len(l.eps_src) for l in lats])])
        self.keep_host = [l.keep_mask(a) for l, a in zip(lats, num_alis)]
        host = {
            "seq_state_off": sbase.astype(np.int32),
            "lvl_base": lvl_base.astype(np.int32),
            "level_off": np.concatenate([l.level_off + np.int32(sbase[i]) for i, l in enumerate(lats)]).astype(np.int32),
            "num_frames": np.array(self.num_frames_host, np.int32),
            "out_off": _cat_off([l.out_off for l in lats]),
            "out_dst": np.concatenate([l.out_dst + np.int32(sbase[i]) for i, l in enumerate(lats)]),
            "out_tid": np.concatenate([l.out_tid for l in lats]),
            "out_gc": np.concatenate([l.out_gc for l in lats]),
            "in_off": _cat_off([l.in_off for l in lats]),
            "in_src": np.concatenate([l.in_src + np.int32(sbase[i]) for i, l in enumerate(lats)]),
            "in_tid": np.concatenate([l.in_tid for l in lats]),
            "in_gc": np.concatenate([l.in_gc for l in lats]),
            # eps_off is indexed with lvl_base[b] + t like level_off (T+2 entries per lattice)
            "eps_off": np.concatenate([l.eps_off + ebase[i] for i, l in enumerate(lats)]).astype(np.int32),
            "eps_src": np.concatenate([l.eps_src + sbase[i] for i, l in enumerate(lats)] + [np.zeros(1, np.int32)]).astype(np.int32),
            "eps_dst": np.concatenate([l.eps_dst + sbase[i] for i, l in enumerate(lats)] + [np.zeros(1, np.int32)]).astype(np.int32),
            "eps_gc": np.concatenate([l.eps_gc for l in lats] + [np.zeros(1, np.float32)]),
            "final_cost": np.concatenate([l.final_cost for l in lats]),
            "state_time": np.concatenate([l.state_time for l in lats]),
            "tid2pdf": np.asarray(tid2pdf, np.int32),
            "num_ali": np.concatenate([np.asarray(a, np.int32) for a in num_alis]),
            "frame_base": np.concatenate([[0], np.cumsum(self.num_frames_host)]).astype(np.int32),
            "keep": np.concatenate(self.keep_host).astype(np.uint8),
        }
        self.mpe = mpe
        if mpe is not None:
            criterion, tid2phone, silence = mpe
            accs = [l.frame_acc(a, tid2pdf, tid2phone, criterion, silence) for l, a in zip(lats, num_alis)]
            host["acc_in"] = np.concatenate([a[0] for a in accs] + [np.zeros(1, np.uint8)])
            host["acc_out"] = np.concatenate([a[1] for a in accs] + [np.zeros(1, np.uint8)])
        self._dev, self._keep = _upload(host, device)
        self.acc_in = self._dev.pop("acc_in", None)
        self.acc_out = self._dev.pop("acc_out", None)
        self.struct = _lib.LatBatch()
        self.struct.n_seq = self.n_seq
        for k, t in self._dev.items():
            setattr(self.struct, k, t.data_ptr())
        self.h2d_bytes = sum(v.nbytes for v in host.values())


class SyntheticLatticeProvider(object):
    """Duck-types the reference's ``asr_decoder`` (bin/train_se.py:179-181): ``decode(loglikes)``
    returns ``{"lattice": Lattice}``.  Lattices are supplied per utterance (synthetic,
    fixed topology -- BASELINE config 3); the acoustic costs are never materialised: the
    kernel gathers them from the loglike matrix, which is what the decoder would have
    written into the arcs."""

    def __init__(self, lattices=None):
        self._queue = list(lattices or [])

    def push(self, lattice):
        self._queue.append(lattice)

    def decode(self, loglikes):
        if not self._queue:
            raise RuntimeError("SyntheticLatticeProvider: no lattice queued for this utterance")
        lat = self._queue.pop(0)
        if lat.num_frames != loglikes.shape[0]:
            raise RuntimeError("lattice has %d frames, loglikes %d" % (lat.num_frames, loglikes.shape[0]))
        return {"lattice": lat}
