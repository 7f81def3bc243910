"""Lattice forward-backward + MMI posterior merge: float64 CPU restatement.
TEST INFRASTRUCTURE ONLY.

Restates what ops/ops.py:52-75 (MMIFunction) reaches through PyKaldi:
top_sort -> ScaleLattice(lm=1.0, ac=0.2) (ops/ops.py:57-59) ->
lattice_forward_backward_mmi(trans_model, lat, num_ali, drop_frames=True,
convert_to_pdf_ids=False, cancel=True) (ops/ops.py:60) ->
Posterior.to_pdf_matrix (ops/ops.py:61-62); backward = -post_mat (ops/ops.py:68-75).
The arithmetic lives in Kaldi (src/lat/lattice-functions.cc, src/hmm/posterior.cc),
not vendored / not pinned by the reference: parity unpinned by the reference;
pinned here by brute-force enumeration and autograd (tests/test_oracle_lattice.py).
Formulas: SURVEY.md Appendix B.

A lattice is a dict of numpy arrays: num_states, src[A], dst[A], tid[A]
(0 = epsilon), graph_cost[A], final_cost[S] (+inf = not final); start state 0;
arcs sorted by src and src < dst or same-level epsilon in topological order.
The acoustic cost of a non-epsilon arc leaving a state of time t is
-loglikes[t, pdf(tid)] (what the decoder would have put there).
"""
import numpy as np


def lattice_state_times(lat):
    """Kaldi LatticeStateTimes: time[dst] = time[src] + (ilabel != 0)."""
    S = int(lat["num_states"])
    times = np.full(S, -1, np.int32)
    times[0] = 0
    for s, d, l in zip(lat["src"], lat["dst"], lat["tid"]):
        assert times[s] >= 0
        nt = times[s] + (1 if l != 0 else 0)
        if times[d] < 0:
            times[d] = nt
        else:
            assert times[d] == nt
    return times


def lattice_fb_mmi(loglikes, lat, tid2pdf, num_ali, lm_scale=1.0, ac_scale=0.2):
    """Returns (tot_like, post_mat [T,N] float64, drop_mask [T] bool, state_times).

    post_mat = pdf-level (numerator - denominator) posteriors after cancel/drop.
    torch gradient of MMIFunction = -post_mat.
    """
    ll = np.asarray(loglikes, np.float64)
    T, N = ll.shape
    S = int(lat["num_states"])
    src = np.asarray(lat["src"], np.int64)
    dst = np.asarray(lat["dst"], np.int64)
    tid = np.asarray(lat["tid"], np.int64)
    gc = np.asarray(lat["graph_cost"], np.float64)
    fin = np.asarray(lat["final_cost"], np.float64)
    tid2pdf = np.asarray(tid2pdf, np.int64)
    times = lattice_state_times(lat).astype(np.int64)
    A = len(src)

    like = np.empty(A)
    for k in range(A):
        if tid[k] != 0:
            like[k] = -lm_scale * gc[k] + ac_scale * ll[times[src[k]], tid2pdf[tid[k]]]
        else:
            like[k] = -lm_scale * gc[k]

    alpha = np.full(S, -np.inf)
    alpha[0] = 0.0
    for k in range(A):
        alpha[dst[k]] = np.logaddexp(alpha[dst[k]], alpha[src[k]] + like[k])
    fmask = np.isfinite(fin)
    assert (times[fmask] == T).all()
    tot = -np.inf
    for s in np.nonzero(fmask)[0]:
        tot = np.logaddexp(tot, alpha[s] - lm_scale * fin[s])
    beta = np.where(fmask, -lm_scale * fin, -np.inf)
    den = [dict() for _ in range(T)]
    for k in range(A - 1, -1, -1):
        ab = beta[dst[k]] + like[k]
        beta[src[k]] = np.logaddexp(beta[src[k]], ab)
        if tid[k] != 0:
            t = times[src[k]]
            den[t][tid[k]] = den[t].get(tid[k], 0.0) + np.exp(alpha[src[k]] + ab - tot)

    post = np.zeros((T, N))
    drop = np.zeros(T, bool)
    for t in range(T):
        nt = int(num_ali[t])
        if nt not in den[t]:           # disjoint -> frame cleared (drop_frames=True)
            drop[t] = True
            continue
        merged = {k: -v for k, v in den[t].items()}
        merged[nt] = merged.get(nt, 0.0) + 1.0
        for k, v in merged.items():
            if v != 0.0:               # cancel=True: exact zeros removed
                post[t, tid2pdf[k]] += v
    return float(tot), post, drop, times.astype(np.int32)


def mpe_frame_acc(tid, ref_tid, tid2pdf, tid2phone, criterion, silence_phones, one_silence_class=True):
    """Per-arc frame accuracy of Kaldi's LatticeForwardBackwardMpeVariants (lat/lattice-functions.cc):
    smbr compares pdfs, mpfe phones; with one_silence_class (what ops/ops.py:138 passes) an arc also counts
    as correct when its phone and the reference phone are both silence phones."""
    phone, ref_phone = int(tid2phone[tid]), int(tid2phone[ref_tid])
    phone_is_sil = phone in silence_phones
    both_sil = phone_is_sil and (ref_phone in silence_phones)
    if criterion == "mpfe":
        same = phone == ref_phone
    elif criterion == "smbr":
        same = int(tid2pdf[tid]) == int(tid2pdf[ref_tid])
    else:
        raise ValueError("criterion must be 'smbr' or 'mpfe'")
    if one_silence_class:
        return 1.0 if (same or both_sil) else 0.0
    return 1.0 if (same and not phone_is_sil) else 0.0


def lattice_fb_mpe(loglikes, lat, tid2pdf, tid2phone, num_ali, criterion, silence_phones,
                   one_silence_class=True, lm_scale=1.0, ac_scale=1.0):
    """sMBR / MPFE: restates what ops/ops.py:130-147 (sMBRFunction) reaches through PyKaldi:
    lattice_forward_backward_mpe_variants(trans_model, silence_phones, lattice, trans_ids, criterion, True)
    followed by Posterior.to_pdf_matrix.  NOTE the reference does not call lattice_scale on this path
    (ops/ops.py:133-143), so graph and acoustic scores both enter with scale 1.
    Two passes in the order of Kaldi's LatticeForwardBackwardMpeVariants: log-domain alpha/beta, then the
    expected-accuracy recursions alpha_smbr / beta_smbr; arc posterior * (accuracy through the arc - expected
    accuracy) is pushed to (t, tid) and summed per pdf.
    Returns (tot_forward_score = expected frame accuracy, post_mat [T,N] float64, tot_like).
    torch gradient of sMBRFunction = -post_mat (ops/ops.py:149-156)."""
    ll = np.asarray(loglikes, np.float64)
    T, N = ll.shape
    S = int(lat["num_states"])
    src = np.asarray(lat["src"], np.int64)
    dst = np.asarray(lat["dst"], np.int64)
    tid = np.asarray(lat["tid"], np.int64)
    gc = np.asarray(lat["graph_cost"], np.float64)
    fin = np.asarray(lat["final_cost"], np.float64)
    tid2pdf = np.asarray(tid2pdf, np.int64)
    silence_phones = set(int(p) for p in silence_phones)
    times = lattice_state_times(lat).astype(np.int64)
    A = len(src)
    assert len(num_ali) == T
    like = np.empty(A)
    acc = np.zeros(A)
    for k in range(A):
        if tid[k] != 0:
            t = times[src[k]]
            like[k] = -lm_scale * gc[k] + ac_scale * ll[t, tid2pdf[tid[k]]]
            acc[k] = mpe_frame_acc(tid[k], int(num_ali[t]), tid2pdf, tid2phone, criterion, silence_phones,
                                   one_silence_class)
        else:
            like[k] = -lm_scale * gc[k]
    fmask = np.isfinite(fin)
    # first pass
    alpha = np.full(S, -np.inf)
    alpha[0] = 0.0
    for k in range(A):
        alpha[dst[k]] = np.logaddexp(alpha[dst[k]], alpha[src[k]] + like[k])
    tot = -np.inf
    for s in np.nonzero(fmask)[0]:
        tot = np.logaddexp(tot, alpha[s] - lm_scale * fin[s])
    beta = np.where(fmask, -lm_scale * fin, -np.inf)
    for k in range(A - 1, -1, -1):
        beta[src[k]] = np.logaddexp(beta[src[k]], beta[dst[k]] + like[k])
    # second pass forward
    alpha_s = np.zeros(S)
    for k in range(A):
        sc = np.exp(alpha[src[k]] + like[k] - alpha[dst[k]])
        alpha_s[dst[k]] += sc * (alpha_s[src[k]] + acc[k])
    tot_score = 0.0
    for s in np.nonzero(fmask)[0]:
        tot_score += np.exp(alpha[s] - lm_scale * fin[s] - tot) * alpha_s[s]
    # second pass backward + posteriors
    beta_s = np.zeros(S)
    post = np.zeros((T, N))
    for k in range(A - 1, -1, -1):
        ab = beta[dst[k]] + like[k]
        sc = np.exp(ab - beta[src[k]])
        beta_s[src[k]] += sc * (beta_s[dst[k]] + acc[k])
    for k in range(A):
        if tid[k] != 0:
            p = np.exp(alpha[src[k]] + beta[dst[k]] + like[k] - tot)
            diff = alpha_s[src[k]] + acc[k] + beta_s[dst[k]] - tot_score
            post[times[src[k]], tid2pdf[tid[k]]] += p * diff
    return float(tot_score), post, float(tot)
