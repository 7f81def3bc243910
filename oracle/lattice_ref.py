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
