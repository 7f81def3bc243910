"""LF-MMI (chain) objective: float64 CPU restatement.  TEST INFRASTRUCTURE ONLY.

Restates what the reference reaches through PyKaldi at ops/ops.py:252-273
(``kaldi_chain.compute_chain_objf_and_deriv``) and bin/train_chain.py:167,202
(``DenominatorGraph(den_fst, num_pdfs)``).  The arithmetic lives in Kaldi
(src/chain/chain-den-graph.cc, chain-denominator.cc, chain-numerator.cc,
chain-training.cc), which is NOT vendored in /root/reference and not pinned by
its Dockerfile (docker/Dockerfile:57-64): parity unpinned by the reference.  The
formulas followed are SURVEY.md Appendix C; they are pinned by
tests/test_oracle_chain.py (brute force, autograd, scaled-vs-log agreement).

An FST here is a dict of numpy arrays:
  num_states, start, src[A], dst[A], ilabel[A] (pdf+1, epsilon-free),
  weight[A] (cost = -log prob), final[S] (cost, +inf = not final)
with arcs sorted by src (OpenFst arc order).
"""
import numpy as np


# ----------------------------------------------------------------------------
# DenominatorGraph  (Kaldi chain-den-graph.cc: SetTransitions / SetInitialProbs)
# ----------------------------------------------------------------------------
def den_graph_from_fst(fst, num_pdfs):
    """Return the arrays Kaldi's DenominatorGraph holds.

    forward_transitions: for state i, arcs (prob, pdf, dst) in FST arc order
    backward_transitions: for state j, arcs (prob, pdf, src) in order of the
      global arc enumeration (state-major, arc-minor), as Kaldi pushes them.
    initial_probs: 100-step propagation, averaged, normalised (double -> f32).
    """
    S = int(fst["num_states"])
    src = np.asarray(fst["src"], np.int64)
    dst = np.asarray(fst["dst"], np.int64)
    pdf = np.asarray(fst["ilabel"], np.int64) - 1
    assert (pdf >= 0).all() and (pdf < num_pdfs).all(), "den fst must be epsilon-free"
    assert (np.diff(src) >= 0).all(), "arcs must be sorted by source state"
    prob = np.exp(-np.asarray(fst["weight"], np.float64)).astype(np.float32)

    fwd_off = np.zeros(S + 1, np.int32)
    np.add.at(fwd_off, src + 1, 1)
    fwd_off = np.cumsum(fwd_off).astype(np.int32)
    order_b = np.argsort(dst, kind="stable")
    bwd_off = np.zeros(S + 1, np.int32)
    np.add.at(bwd_off, dst + 1, 1)
    bwd_off = np.cumsum(bwd_off).astype(np.int32)

    g = {
        "num_states": S,
        "num_pdfs": int(num_pdfs),
        "fwd_off": fwd_off,
        "fwd_prob": prob.copy(),
        "fwd_pdf": pdf.astype(np.int32),
        "fwd_state": dst.astype(np.int32),
        "bwd_off": bwd_off,
        "bwd_prob": prob[order_b].copy(),
        "bwd_pdf": pdf[order_b].astype(np.int32),
        "bwd_state": src[order_b].astype(np.int32),
    }

    # initial probs (double)
    final_p = np.exp(-np.asarray(fst["final"], np.float64))  # exp(-inf)=0
    probd = np.exp(-np.asarray(fst["weight"], np.float64))
    tot = final_p.copy()
    np.add.at(tot, src, probd)
    norm = 1.0 / tot
    cur = np.zeros(S, np.float64)
    cur[int(fst["start"])] = 1.0
    avg = np.zeros(S, np.float64)
    for _ in range(100):
        avg += cur / 100.0
        nxt = np.zeros(S, np.float64)
        np.add.at(nxt, dst, cur[src] * norm[src] * probd)
        nxt /= nxt.sum()
        cur = nxt
    g["initial_probs"] = avg.astype(np.float32)
    return g


# ----------------------------------------------------------------------------
# Denominator forward-backward
# ----------------------------------------------------------------------------
def den_fb_scaled(loglikes, g, leaky=1e-4, dtype=np.float64):
    """Kaldi's scaled-probability recursion, in `dtype`.

    loglikes: [T, N].  Returns (logZ_den, gamma_den [T,N], ok).
    """
    ll = np.asarray(loglikes, np.float64)
    T, N = ll.shape
    S = g["num_states"]
    e = np.exp(np.clip(ll, -30.0, 30.0)).astype(dtype)
    init = g["initial_probs"].astype(dtype)
    # arc list in forward order
    fsrc = np.repeat(np.arange(S), np.diff(g["fwd_off"]))
    fdst = g["fwd_state"].astype(np.int64)
    fpdf = g["fwd_pdf"].astype(np.int64)
    fw = g["fwd_prob"].astype(dtype)

    alpha_dash = np.zeros((T + 1, S), dtype)
    asum = np.zeros(T + 1, dtype)
    a = init.copy()
    asum[0] = a.sum()
    alpha_dash[0] = a + leaky * asum[0] * init
    for t in range(1, T + 1):
        contrib = alpha_dash[t - 1, fsrc] * fw * e[t - 1, fpdf]
        a = np.zeros(S, dtype)
        np.add.at(a, fdst, contrib)
        a = a / asum[t - 1]
        asum[t] = a.sum()
        alpha_dash[t] = a + leaky * asum[t] * init
    totp = alpha_dash[T].sum()
    logZ = float(np.log(np.float64(totp)) + np.log(asum[:T].astype(np.float64)).sum())

    gamma = np.zeros((T, N), dtype)
    beta_dash = np.full(S, dtype(1.0) / totp, dtype)
    beta = beta_dash + leaky * (beta_dash * init).sum()
    for t in range(T - 1, -1, -1):
        x = fw * e[t, fpdf] * beta[fdst] / asum[t]
        bd = np.zeros(S, dtype)
        np.add.at(bd, fsrc, x)
        np.add.at(gamma[t], fpdf, alpha_dash[t, fsrc] * x)
        beta_dash = bd
        beta = beta_dash + leaky * (beta_dash * init).sum()
    ok = bool(np.isfinite(logZ))
    return logZ, gamma.astype(np.float64), ok


def den_fb_log(loglikes, g, leaky=1e-4):
    """Independent pure log-domain formulation (no per-frame scaling).

    The leaky-HMM step is written as a rank-one update in log space.
    Returns (logZ_den, gamma_den).  Used only to cross-check den_fb_scaled.
    """
    ll = np.clip(np.asarray(loglikes, np.float64), -30.0, 30.0)
    T, N = ll.shape
    S = g["num_states"]
    with np.errstate(divide="ignore"):
        linit = np.log(g["initial_probs"].astype(np.float64))
        lw = np.log(g["fwd_prob"].astype(np.float64))
    fsrc = np.repeat(np.arange(S), np.diff(g["fwd_off"]))
    fdst = g["fwd_state"].astype(np.int64)
    fpdf = g["fwd_pdf"].astype(np.int64)
    lleaky = np.log(leaky) if leaky > 0 else -np.inf

    def lse(v):
        m = np.max(v)
        if not np.isfinite(m):
            return -np.inf
        return m + np.log(np.exp(v - m).sum())

    def scatter_lse(idx, val, n):
        m = np.full(n, -np.inf)
        np.maximum.at(m, idx, val)
        ms = np.where(np.isfinite(m), m, 0.0)
        acc = np.zeros(n)
        np.add.at(acc, idx, np.exp(val - ms[idx]))
        with np.errstate(divide="ignore"):
            return ms + np.log(acc)

    def dash(la):
        return np.logaddexp(la, lleaky + lse(la) + linit)

    la_d = np.zeros((T + 1, S))
    la_d[0] = dash(linit)
    for t in range(1, T + 1):
        la = scatter_lse(fdst, la_d[t - 1, fsrc] + lw + ll[t - 1, fpdf], S)
        la_d[t] = dash(la)
    logZ = lse(la_d[T])

    gamma = np.zeros((T, N))
    lbd = np.zeros(S)  # beta'(T) = 1 (unnormalised); divide by Z at the end
    lb = np.logaddexp(lbd, lleaky + lse(lbd + linit))
    for t in range(T - 1, -1, -1):
        x = lw + ll[t, fpdf] + lb[fdst]
        np.add.at(gamma[t], fpdf, np.exp(la_d[t, fsrc] + x - logZ))
        lbd = scatter_lse(fsrc, x, S)
        lb = np.logaddexp(lbd, lleaky + lse(lbd + linit))
    return float(logZ), gamma


# ----------------------------------------------------------------------------
# Numerator forward-backward (Kaldi chain-numerator.cc: log domain, double)
# ----------------------------------------------------------------------------
def fst_state_times(fst):
    """Time stamp of every state of an epsilon-free, topologically sorted FST."""
    S = int(fst["num_states"])
    times = np.full(S, -1, np.int32)
    times[int(fst["start"])] = 0
    for s, d in zip(fst["src"], fst["dst"]):
        assert times[s] >= 0, "fst not topologically sorted / not connected"
        if times[d] < 0:
            times[d] = times[s] + 1
        else:
            assert times[d] == times[s] + 1, "paths of different length reach one state"
    return times


def num_fb_log(loglikes, fst):
    """Log-domain forward-backward over the supervision FST.

    Returns (logZ_num, gamma_num [T,N]).  Arc score = loglikes[time(src), pdf] - weight.
    """
    ll = np.asarray(loglikes, np.float64)
    T, N = ll.shape
    S = int(fst["num_states"])
    src = np.asarray(fst["src"], np.int64)
    dst = np.asarray(fst["dst"], np.int64)
    pdf = np.asarray(fst["ilabel"], np.int64) - 1
    w = np.asarray(fst["weight"], np.float64)
    fin = np.asarray(fst["final"], np.float64)
    times = fst_state_times(fst).astype(np.int64)
    la = np.full(S, -np.inf)
    la[int(fst["start"])] = 0.0
    A = len(src)
    for k in range(A):  # arcs sorted by src; src order is topological
        la[dst[k]] = np.logaddexp(la[dst[k]], la[src[k]] + ll[times[src[k]], pdf[k]] - w[k])
    fmask = np.isfinite(fin)
    assert (times[fmask] == T).all(), "final states must have time T"
    logZ = -np.inf
    for s in np.nonzero(fmask)[0]:
        logZ = np.logaddexp(logZ, la[s] - fin[s])
    lb = np.where(fmask, -fin, -np.inf)
    gamma = np.zeros((T, N))
    for k in range(A - 1, -1, -1):
        sc = ll[times[src[k]], pdf[k]] - w[k] + lb[dst[k]]
        lb[src[k]] = np.logaddexp(lb[src[k]], sc)
        gamma[times[src[k]], pdf[k]] += np.exp(la[src[k]] + sc - logZ)
    return float(logZ), gamma


# ----------------------------------------------------------------------------
# compute_chain_objf_and_deriv  (Kaldi chain-training.cc) as used at ops/ops.py:265-273
# ----------------------------------------------------------------------------
def chain_objf_and_deriv(loglikes, den, sup_fst, weight=1.0, leaky=1e-4, xent_regularize=0.0):
    """Returns (objf, grad_kaldi, grad_xent) with grad_kaldi already including
    xent_regularize*grad_xent (ops/ops.py:267).  torch grad = -grad_kaldi (ops/ops.py:275-280)."""
    T = loglikes.shape[0]
    logZ_den, g_den, ok = den_fb_scaled(loglikes, den, leaky)
    logZ_num, g_num = num_fb_log(loglikes, sup_fst)
    objf = weight * (logZ_num - logZ_den)
    grad = weight * (g_num - g_den)
    grad_xent = weight * g_num
    if not ok or not np.isfinite(objf):
        objf = -10.0 * weight * T
        grad = np.zeros_like(grad)
        grad_xent = np.zeros_like(grad)
    grad = grad + xent_regularize * grad_xent
    return float(objf), grad, grad_xent
