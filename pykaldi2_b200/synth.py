"""Seeded synthetic inputs for the hot path.

No zip/wav I/O, no Kaldi assets: waveforms, labels, denominator FSTs, numerator
(supervision) FSTs and decoding lattices are generated here, deterministically,
in the same array formats the host classes in ``graphs.py`` consume.  These stand
in for what the reference reads from disk / gets from PyKaldi
(bin/train_chain.py:167-202,262-272; bin/train_se.py:145-181; ops/ops.py:55).
"""
import numpy as np

SAMPLE_RATE = 16000


def make_durations(n, rng, lo=1.5, hi=30.0):
    """LibriSpeech-shaped utterance durations in seconds (mean ~12.3 s)."""
    return np.clip(rng.gamma(6.0, 2.05, size=n), lo, hi)


def make_waveforms(durs, rng, amp=0.05):
    """float32 waveforms in [-1,1) at 16 kHz, one per duration."""
    return [(amp * rng.standard_normal(int(round(d * SAMPLE_RATE)))).astype(np.float32) for d in durs]


def make_den_fst(num_states=8192, num_pdfs=5768, mean_extra=7, seed=1234):
    """Epsilon-free stochastic denominator FST.

    start 0; backbone arc i -> (i+1) mod S plus U{mean_extra-4 .. mean_extra+4}
    extra arcs to uniform random states (mean out-degree 1+mean_extra);
    pdf ~ U[0,N); outgoing probs of a state = softmax(N(0,1)); no final weights.
    Returns the FST dict format used by oracle/chain_ref.py and graphs.DenGraph.
    """
    rng = np.random.default_rng(seed)
    S = int(num_states)
    lo, hi = max(0, mean_extra - 4), mean_extra + 4
    extra = rng.integers(lo, hi + 1, size=S)
    deg = extra + 1
    A = int(deg.sum())
    off = np.concatenate([[0], np.cumsum(deg)])
    src = np.repeat(np.arange(S, dtype=np.int32), deg)
    dst = rng.integers(0, S, size=A).astype(np.int32)
    dst[off[:-1]] = (np.arange(S) + 1) % S          # first arc of each state = backbone
    pdf = rng.integers(0, num_pdfs, size=A).astype(np.int32)
    z = rng.standard_normal(A)
    m = np.maximum.reduceat(z, off[:-1])
    ez = np.exp(z - np.repeat(m, deg))
    p = ez / np.repeat(np.add.reduceat(ez, off[:-1]), deg)
    return {
        "num_states": S, "start": 0,
        "src": src, "dst": dst, "ilabel": pdf + 1,
        "weight": (-np.log(p)).astype(np.float32),
        "final": np.full(S, np.inf, np.float32),
    }


def make_supervision_fst(T, num_pdfs, rng, slack=2, min_dur=2, max_dur=8):
    """Time-constrained numerator FST for one utterance (T output frames).

    A random segmentation into pdf-labelled segments; each segment boundary may
    move by +-slack frames.  State = (time t, segment k); arcs (t,k)->(t+1,k)
    labelled pdf_k (stay) and (t,k)->(t+1,k+1) labelled pdf_{k+1} (advance).
    States are numbered in (time, segment) order, hence topologically sorted with
    non-decreasing time stamps; arc weights 0 (bin/train_chain.py:271-272 builds
    the supervision without a normalization FST); the final state has time T.
    """
    T = int(T)
    durs = []
    tot = 0
    while tot < T:
        d = int(rng.integers(min_dur, max_dur + 1))
        durs.append(d)
        tot += d
    durs[-1] -= tot - T
    if durs[-1] <= 0:
        durs.pop()
        durs[-1] += T - sum(durs)
    K = len(durs)
    pdfs = rng.integers(0, num_pdfs, size=K)
    return segments_to_supervision_fst(pdfs, durs, T, slack)


def segments_to_supervision_fst(pdfs, durs, T, slack=2):
    """Numerator FST of a segmentation: segment k carries pdf ``pdfs[k]`` for nominally ``durs[k]`` frames
    (sum(durs) == T); every boundary may move by +-slack frames.  See make_supervision_fst for the state layout."""
    T = int(T)
    K = len(durs)
    pdfs = np.asarray(pdfs)
    ends = np.cumsum(durs)                  # nominal end frame (exclusive) of segment k
    # segment k may be active at frame t iff start_k - slack <= t < end_k + slack
    starts = np.concatenate([[0], ends[:-1]])
    lo = np.maximum(starts - slack, 0)
    lo[0] = 0
    hi = np.minimum(ends + slack, T)
    hi[-1] = T
    # make windows consistent: a path must be able to advance one segment per frame at most
    for k in range(1, K):
        lo[k] = max(lo[k], lo[k - 1] + 1)
    for k in range(K - 2, -1, -1):
        hi[k] = min(hi[k], hi[k + 1] - 1)
    # state (t,k): "about to emit frame t while in segment k" for lo[k] <= t < hi[k]; plus final (T, K-1)
    sid = {}
    for t in range(T):
        for k in range(K):
            if lo[k] <= t < hi[k]:
                sid[(t, k)] = len(sid)
    final_id = len(sid)
    src, dst, lab = [], [], []
    for (t, k), s in sid.items():           # insertion order = (t,k) order = state order
        if t + 1 == T:
            if k == K - 1:
                src.append(s); dst.append(final_id); lab.append(pdfs[k] + 1)
            continue
        if (t + 1, k) in sid:
            src.append(s); dst.append(sid[(t + 1, k)]); lab.append(pdfs[k] + 1)
        if (t + 1, k + 1) in sid:
            src.append(s); dst.append(sid[(t + 1, k + 1)]); lab.append(pdfs[k] + 1)
    S = final_id + 1
    fst = {
        "num_states": S, "start": 0,
        "src": np.asarray(src, np.int32), "dst": np.asarray(dst, np.int32),
        "ilabel": np.asarray(lab, np.int32),
        "weight": np.zeros(len(src), np.float32),
        "final": np.full(S, np.inf, np.float32),
    }
    fst["final"][final_id] = 0.0
    fst["state_times"] = np.asarray([t for (t, k) in sid] + [T], np.int32)
    return _trim(fst)


def alignment_to_supervision_fst(pdf_ali, factor=3, shift=0, slack=2, n_out=None):
    """Numerator FST from a frame-level pdf alignment: the alignment is subsampled like the features
    (frames shift, shift + factor, ...: bin/train_chain.py:251-255), run-length encoded into segments, and every
    segment boundary may move by +-slack output frames.  A stand-in for the reference's
    alignment -> phone/durations -> proto-supervision (tolerance) -> pdf FST chain (bin/train_chain.py:262-272), which
    needs Kaldi's tree and topology: the time tolerance is applied at the pdf level instead of the phone level."""
    ali = np.asarray(pdf_ali).reshape(-1)[shift::factor]
    if n_out is not None:
        ali = ali[:n_out]
        if len(ali) < n_out:                # features may be a frame or two longer than the labels
            ali = np.concatenate([ali, np.full(n_out - len(ali), ali[-1], ali.dtype)])
    if len(ali) == 0:
        raise ValueError("empty alignment")
    change = np.flatnonzero(np.diff(ali)) + 1
    starts = np.concatenate([[0], change])
    durs = np.diff(np.concatenate([starts, [len(ali)]])).tolist()
    return segments_to_supervision_fst(ali[starts].astype(np.int64), durs, len(ali), slack)


def _trim(fst):
    """Remove states that are not both accessible and co-accessible (keeps order)."""
    S = fst["num_states"]
    src, dst = fst["src"], fst["dst"]
    acc = np.zeros(S, bool); acc[fst["start"]] = True
    for s, d in zip(src, dst):              # arcs sorted by src, topological
        if acc[s]:
            acc[d] = True
    co = np.isfinite(fst["final"]).copy()
    for s, d in zip(src[::-1], dst[::-1]):
        if co[d]:
            co[s] = True
    keep = acc & co
    assert keep[fst["start"]], "supervision has no successful path"
    newid = np.cumsum(keep) - 1
    ak = keep[src] & keep[dst]
    out = {
        "num_states": int(keep.sum()), "start": int(newid[fst["start"]]),
        "src": newid[src[ak]].astype(np.int32), "dst": newid[dst[ak]].astype(np.int32),
        "ilabel": fst["ilabel"][ak], "weight": fst["weight"][ak],
        "final": fst["final"][keep],
    }
    if "state_times" in fst:
        out["state_times"] = fst["state_times"][keep]
    return out


def make_lattice(T, num_pdfs, rng, num_ali=None, kmin=32, kmax=96, dmin=2, dmax=6,
                 ali_drop=0.05, eps_frac=0.0, num_tids=None):
    """Frame-layered decoding lattice.

    K_t ~ U{kmin..kmax} states at each time 1..T (one start state at t=0), each
    state has U{dmin..dmax} arcs to states at t+1; tid ~ U{1..2N} with
    pdf(tid) = (tid-1)//2; graph cost ~ U(0,8); all states at t=T are final with
    cost U(0,2).  The reference alignment num_ali[t] is carried by one arc of
    frame t on (1-ali_drop) of the frames; the others deliberately lack it
    (exercises drop_frames).  eps_frac>0 adds epsilon (tid 0) arcs between
    same-time states (lower index -> higher index).
    ``num_tids``: draw the arcs' transition ids from 1..num_tids (a real transition model's id range) instead of
    1..2N; the returned tid2pdf is the synthetic two-ids-per-pdf map either way.
    Returns (lattice dict, tid2pdf int32 [2N+1], num_ali int32 [T]).
    """
    T = int(T)
    N = int(num_pdfs)
    n_tid = 2 * N if num_tids is None else int(num_tids)
    tid2pdf = np.concatenate([[-1], np.repeat(np.arange(N), 2)]).astype(np.int32)
    if num_ali is None:
        num_ali = rng.integers(1, n_tid + 1, size=T).astype(np.int32)
    src_l, dst_l, tid_l, gc_l = [], [], [], []
    base, ns = 0, 1                      # first state id and number of states of the current level
    for t in range(T):
        deg = rng.integers(dmin, dmax + 1, size=ns)
        s = np.repeat(np.arange(ns), deg) + base
        a = int(deg.sum())
        cand = int(rng.integers(kmin, kmax + 1))
        d = rng.integers(0, cand, size=a)
        # decoder lattices are connected: only candidate states some arc reaches exist
        hit = np.zeros(cand, bool)
        hit[d] = True
        newid = np.cumsum(hit) - 1
        nxt = int(hit.sum())
        d = newid[d] + base + ns
        tid = rng.integers(1, n_tid + 1, size=a)
        tid[tid == num_ali[t]] = (num_ali[t] % n_tid) + 1     # never hit the alignment by accident
        if rng.random() >= ali_drop:
            tid[rng.integers(0, a)] = num_ali[t]
        gc = rng.uniform(0.0, 8.0, size=a)
        if eps_frac > 0 and ns > 1:
            ne = max(1, int(eps_frac * a))
            es = rng.integers(0, ns - 1, size=ne)
            ed = es + 1 + rng.integers(0, ns - 1 - es)
            s = np.concatenate([s, es + base])
            d = np.concatenate([d, ed + base])
            tid = np.concatenate([tid, np.zeros(ne, np.int64)])
            gc = np.concatenate([gc, rng.uniform(0.0, 2.0, size=ne)])
            o = np.argsort(s, kind="stable")
            s, d, tid, gc = s[o], d[o], tid[o], gc[o]
        src_l.append(s); dst_l.append(d); tid_l.append(tid); gc_l.append(gc)
        base, ns = base + ns, nxt
    S = base + ns
    final = np.full(S, np.inf, np.float32)
    final[base:] = rng.uniform(0.0, 2.0, size=ns).astype(np.float32)
    lat = {
        "num_states": int(S),
        "src": np.concatenate(src_l).astype(np.int32),
        "dst": np.concatenate(dst_l).astype(np.int32),
        "tid": np.concatenate(tid_l).astype(np.int32),
        "graph_cost": np.concatenate(gc_l).astype(np.float32),
        "final_cost": final,
    }
    return lat, tid2pdf, np.asarray(num_ali, np.int32)


def make_log_prior(num_pdfs, rng):
    occs = rng.dirichlet(np.ones(num_pdfs))
    return np.log(occs / occs.sum()).astype(np.float32)
