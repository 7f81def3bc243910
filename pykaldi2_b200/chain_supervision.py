"""Numerator ("supervision") graphs of LF-MMI from forced alignments -- SURVEY 8f-2, the builder half.

Reference call chain (bin/train_chain.py:184-188, 262-272), all of it inside Kaldi behind PyKaldi:

    phone_ali   = aligner.to_phone_alignment(trans_ids)                  # hmm-utils.cc SplitToPhones
    proto       = alignment_to_proto_supervision(opts, phones, durations)   # chain-supervision.cc
    supervision = proto_supervision_to_supervision(tree, trans_model, proto, convert_to_pdfs=True)

restated here in numpy / Python from Kaldi's published algorithm (src/chain/chain-supervision.cc,
src/hmm/hmm-utils.cc); Kaldi is not vendored in the reference and not available in this sandbox, so parity is
pinned by hand-built cases (tests/test_chain_supervision.py), not by Kaldi output.  Host-side integer / graph work:
nothing here runs on the GPU; the result feeds graphs.Supervision.

What the three steps compute
  1. ``split_to_phones``: cut the transition-id alignment of the ALIGNMENT model into phone segments: a segment ends
     after a transition into the topology's final state (followed, for reordered alignments, by that state's
     self-loops).
  2. ``alignment_to_proto_supervision``: a linear phone acceptor (state i --phone_i--> i+1, self-loop phone_i on
     i+1) plus, per SUBSAMPLED frame, the set of phones allowed there: phone i may occupy the frames
     [start_i - left_tolerance, end_i + right_tolerance), mapped to the subsampled rate by ceil(t / factor).
  3. ``proto_supervision_to_supervision``: expand every phone into the HMM states of its topology entry with the
     pdfs the context-dependency tree gives it in its phone context (H o C o phone acceptor, self-loops added in
     Kaldi's reordered position: one forward transition, then the state's self-loop), replace transition ids by
     pdf-id + 1, and intersect with the per-frame phone constraints (Kaldi's TimeEnforcerFst): states of the result
     are (HMM-state slot, frame) pairs; it is trimmed (fst::Connect) and renumbered breadth first
     (SortBreadthFirstSearch).  Weights are all zero (no normalisation FST on this path: SURVEY Appendix C).
"""
import numpy as np


class SupervisionOptions(object):
    """kaldi_chain.SupervisionOptions (chain-supervision.h): same attribute names and defaults."""

    def __init__(self, left_tolerance=5, right_tolerance=5, frame_subsampling_factor=1, convert_to_pdfs=True):
        self.left_tolerance = left_tolerance
        self.right_tolerance = right_tolerance
        self.frame_subsampling_factor = frame_subsampling_factor
        self.convert_to_pdfs = convert_to_pdfs


def split_to_phones(tm, alignment):
    """SplitToPhones (hmm-utils.cc) on a transition-id alignment.  ``tm``: dict from
    reader.kaldi_io.read_transition_model_text.  Returns [(phone, start_frame, duration), ...] -- what
    MappedAligner.to_phone_alignment hands to the reference trainer.  Handles both the plain order (self-loops
    before the forward transition of a state) and Kaldi's reordered alignments (self-loops after it)."""
    ali = np.asarray(alignment, np.int64)
    n = len(ali)
    if n == 0:
        return []
    if ali.min() < 1 or ali.max() >= len(tm["tid2pdf"]):
        raise ValueError("alignment holds a transition id outside 1..%d" % (len(tm["tid2pdf"]) - 1))
    is_final, is_self, state, phone = (tm["tid_is_final"], tm["tid_is_self_loop"], tm["tid2state"], tm["tid2phone"])
    ends = []
    i = 0
    while i < n:
        t = ali[i]
        if is_final[t]:
            # reordered alignments: the self-loops of the state follow the transition that leaves it
            while i + 1 < n and is_self[ali[i + 1]] and state[ali[i + 1]] == state[t]:
                i += 1
            ends.append(i + 1)
        elif i + 1 == n:
            ends.append(i + 1)                       # truncated alignment: close the last phone at the end
        elif state[t] != state[ali[i + 1]] and phone[t] != phone[ali[i + 1]]:
            ends.append(i + 1)                       # phone change without a final transition (partial phone)
        i += 1
    out, start = [], 0
    for e in ends:
        out.append((int(phone[ali[start]]), start, e - start))
        start = e
    return out


class ProtoSupervision(object):
    """chain::ProtoSupervision: ``allowed_phones[t]`` (sorted unique phones per subsampled frame) and the linear
    phone acceptor as arc arrays (src, dst, label)."""

    def __init__(self, allowed_phones, src, dst, label, num_states):
        self.allowed_phones = allowed_phones
        self.src, self.dst, self.label = src, dst, label
        self.num_states = num_states
        self.phones = [int(l) for s, d, l in zip(src, dst, label) if d == s + 1]


def alignment_to_proto_supervision(opts, phones, durations):
    """AlignmentToProtoSupervision (chain-supervision.cc)."""
    phones = [int(p) for p in phones]
    durations = [int(d) for d in durations]
    if len(phones) != len(durations) or not phones:
        raise ValueError("phones and durations must be non-empty and of equal length")
    if min(phones) <= 0 or min(durations) <= 0:
        raise ValueError("phones and durations must be positive")
    factor = int(opts.frame_subsampling_factor)
    num_frames = sum(durations)
    num_sub = (num_frames + factor - 1) // factor
    allowed = [[] for _ in range(num_sub)]
    src, dst, label = [], [], []
    cur = 0
    for i, (ph, du) in enumerate(zip(phones, durations)):
        t0 = max(0, cur - int(opts.left_tolerance))
        t1 = min(num_frames, cur + du + int(opts.right_tolerance))
        s0, s1 = (t0 + factor - 1) // factor, (t1 + factor - 1) // factor
        for t in range(s0, s1):
            allowed[t].append(ph)
        src += [i, i + 1]; dst += [i + 1, i + 1]; label += [ph, ph]     # arc to the next state, self-loop on it
        cur += du
    allowed = [sorted(set(a)) for a in allowed]
    return ProtoSupervision(allowed, np.asarray(src, np.int32), np.asarray(dst, np.int32), np.asarray(label, np.int32),
                            len(phones) + 1)


def _hmm_slots(tree, tm, phones):
    """Per phone instance, the emitting HMM states of its topology entry as (phone, forward pdf, self-loop pdf,
    has_self_loop, next ...).  Supports the left-to-right topologies Kaldi's recipes use (each emitting state: an
    optional self-loop and one transition to the next state; the last one enters the final state)."""
    N, P = tree.context_width(), tree.central_position()
    slots = []
    for i, ph in enumerate(phones):
        window = []
        for k in range(N):
            j = i + k - P
            window.append(phones[j] if 0 <= j < len(phones) else 0)      # 0 = no phone (utterance boundary)
        topo, classes = tm["topology"].get(ph), tm["pdf_class"].get(ph)
        if topo is None:
            raise ValueError("phone %d has no topology entry in the transition model" % ph)
        hs = 0
        seen = set()
        while len(topo[hs]) > 0:
            if hs in seen:
                raise ValueError("topology of phone %d is not left-to-right" % ph)
            seen.add(hs)
            fwd = [d for d, _ in topo[hs] if d != hs]
            if len(fwd) != 1:
                raise ValueError("topology of phone %d: state %d has %d forward transitions; only left-to-right "
                                 "topologies (one forward transition per state) are supported" % (ph, hs, len(fwd)))
            fc, sc = classes[hs]
            fpdf, spdf = tree.compute(window, fc), tree.compute(window, sc)
            if fpdf is None or spdf is None:
                raise ValueError("the tree has no pdf for phone window %r, state %d" % (window, hs))
            slots.append((ph, int(fpdf), int(spdf), any(d == hs for d, _ in topo[hs])))
            hs = fwd[0]
            if hs >= len(topo):
                break
    return slots


def proto_supervision_to_supervision(tree, tm, proto, convert_to_pdfs=True):
    """ProtoSupervisionToSupervision (chain-supervision.cc) for ``convert_to_pdfs=True`` (the reference's setting,
    bin/train_chain.py:185).  ``tree``: reader.kaldi_io.ContextDependency; ``tm``: the CHAIN transition model (dict of
    read_transition_model_text).  Returns the FST dict graphs.Supervision takes:
    num_states, start, src, dst, ilabel (= pdf + 1), weight (zeros), final (0 at final states, inf elsewhere),
    state_times -- or None when the constraints leave no path (Kaldi: "Supervision FST is empty").

    The transition-id acceptor after AddSelfLoops(reorder=true) + RmEpsilon is, at pdf level, the chain
    q_k --fpdf_k--> q_{k+1}, q_{k+1} --spdf_k--> q_{k+1} over the K HMM-state slots; intersected with the per-frame
    phone constraints its nodes are (q, t).  Every arc goes from frame t to t + 1, so reachability is a T-step
    recursion over boolean vectors of length K + 1 (numpy; a 26 s utterance takes a few milliseconds), and the
    breadth-first numbering of fst::SortBreadthFirstSearch visits the frames in order and, inside a frame, the
    nodes by DECREASING q (parents in decreasing q each emit their forward arc, to q + 1, before their self-loop)."""
    if not convert_to_pdfs:
        raise NotImplementedError("only convert_to_pdfs=True (what the reference trainer sets) is built")
    slots = _hmm_slots(tree, tm, proto.phones)
    K, T = len(slots), len(proto.allowed_phones)
    slot_phone = np.array([sl[0] for sl in slots], np.int64)
    fpdf = np.array([sl[1] for sl in slots], np.int64)
    spdf = np.array([sl[2] for sl in slots], np.int64)
    has_loop = np.array([sl[3] for sl in slots], bool)
    # ok[k, t]: slot k may emit at frame t (its phone is allowed there)
    n_ph = int(max(slot_phone.max(), max((max(a) for a in proto.allowed_phones if a), default=0))) + 1
    allowed = np.zeros((n_ph, T), bool)
    for t, a in enumerate(proto.allowed_phones):
        if a:
            allowed[np.asarray(a, np.int64), t] = True
    ok = allowed[slot_phone]                                   # [K, T]
    fwd_ok = ok                                                # (q, t) -> (q + 1, t + 1) for q < K
    loop_ok = np.zeros((K + 1, T), bool)                       # (q, t) -> (q, t + 1) for q >= 1: self-loop of slot q - 1
    loop_ok[1:] = ok & has_loop[:, None]
    reach = np.zeros((K + 1, T + 1), bool)
    reach[0, 0] = True
    for t in range(T):
        r = reach[:, t]
        nxt = r & loop_ok[:, t]
        nxt[1:] |= r[:-1] & fwd_ok[:, t]
        reach[:, t + 1] = nxt
    if not reach[K, T]:
        return None
    co = np.zeros((K + 1, T + 1), bool)
    co[K, T] = True
    for t in range(T - 1, -1, -1):
        c = co[:, t + 1]
        cur = c & loop_ok[:, t]
        cur[:-1] |= c[1:] & fwd_ok[:, t]
        co[:, t] = cur
    live = reach & co
    # breadth-first ids: frames in order, decreasing q inside a frame
    ids = -np.ones((K + 1, T + 1), np.int64)
    tt, qq = np.nonzero(live.T[:, ::-1])                       # row-major over (t, reversed q)
    qq = K - qq
    ids[qq, tt] = np.arange(len(tt))
    n = len(tt)
    # arcs of node (q, t), t < T: forward (needs slot q allowed at t and (q+1, t+1) live), then self-loop
    src_f = live[:-1, :-1] & fwd_ok & live[1:, 1:]             # [K, T]   q = 0..K-1
    src_l = live[:, :-1] & loop_ok & live[:, 1:]               # [K+1, T]
    qf, tf = np.nonzero(src_f)
    ql, tl = np.nonzero(src_l)
    src = np.concatenate([ids[qf, tf], ids[ql, tl]])
    dst = np.concatenate([ids[qf + 1, tf + 1], ids[ql, tl + 1]])
    lab = np.concatenate([fpdf[qf] + 1, spdf[ql - 1] + 1])
    kind = np.concatenate([np.zeros(len(qf), np.int64), np.ones(len(ql), np.int64)])
    o = np.lexsort((kind, src))                                # by source id, forward arc before self-loop
    final = np.full(n, np.inf, np.float32)
    final[ids[K, T]] = 0.0
    return {
        "num_states": int(n), "start": 0,
        "src": src[o].astype(np.int32), "dst": dst[o].astype(np.int32),
        "ilabel": lab[o].astype(np.int32), "weight": np.zeros(len(src), np.float32),
        "final": final, "state_times": tt.astype(np.int32),
    }


def supervision_from_alignment(opts, ali_tm, chain_tm, tree, trans_ids):
    """The whole reference chain for one utterance (bin/train_chain.py:262-272).  ``ali_tm``: the alignment
    model's transition model (splits the alignment into phones), ``chain_tm`` + ``tree``: the chain model's.
    Returns (fst dict or None, frames_per_sequence)."""
    seg = split_to_phones(ali_tm, trans_ids)
    proto = alignment_to_proto_supervision(opts, [p for p, _, _ in seg], [d for _, _, d in seg])
    return proto_supervision_to_supervision(tree, chain_tm, proto, opts.convert_to_pdfs), len(proto.allowed_phones)
