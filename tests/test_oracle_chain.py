"""Pin the LF-MMI oracle (oracle/chain_ref.py): the reference ships no tests or
golden vectors for this path, so the restatement is pinned by
(1) brute-force path enumeration, (2) autograd through an independent dense
formulation, (3) agreement of two formulations, (4) invariants."""
import itertools

import numpy as np
import pytest
import torch

from oracle import chain_ref
from pykaldi2_b200 import synth


def tiny_den(S, N, seed, deg=3):
    rng = np.random.default_rng(seed)
    src, dst, lab, w = [], [], [], []
    for s in range(S):
        for _ in range(int(rng.integers(1, deg + 1))):
            src.append(s); dst.append(int(rng.integers(0, S)))
            lab.append(int(rng.integers(0, N)) + 1); w.append(float(rng.uniform(0.1, 2.0)))
    fst = {"num_states": S, "start": 0, "src": np.array(src), "dst": np.array(dst),
           "ilabel": np.array(lab), "weight": np.array(w, np.float32),
           "final": np.full(S, np.inf, np.float32)}
    return fst


def brute_den(ll, g):
    """leaky = 0: enumerate all arc paths of length T starting from init."""
    T, N = ll.shape
    S = g["num_states"]
    arcs = [[] for _ in range(S)]
    for i in range(S):
        for k in range(g["fwd_off"][i], g["fwd_off"][i + 1]):
            arcs[i].append((float(g["fwd_prob"][k]), int(g["fwd_pdf"][k]), int(g["fwd_state"][k])))
    Z = 0.0
    num = np.zeros((T, N))

    def rec(t, s, p, pdfs):
        nonlocal Z
        if t == T:
            Z += p
            for tt, q in enumerate(pdfs):
                num[tt, q] += p
            return
        for (w, q, d) in arcs[s]:
            rec(t + 1, d, p * w * np.exp(ll[t, q]), pdfs + [q])

    for s0 in range(S):
        if g["initial_probs"][s0] > 0:
            rec(0, s0, float(g["initial_probs"][s0]), [])
    return np.log(Z), num / Z


def dense_logZ(ll_t, g, leaky):
    """Independent dense-matrix float64 formulation with torch autograd."""
    T, N = ll_t.shape
    S = g["num_states"]
    init = torch.tensor(g["initial_probs"].astype(np.float64))
    fsrc = torch.tensor(np.repeat(np.arange(S), np.diff(g["fwd_off"])))
    fdst = torch.tensor(g["fwd_state"].astype(np.int64))
    fpdf = torch.tensor(g["fwd_pdf"].astype(np.int64))
    fw = torch.tensor(g["fwd_prob"].astype(np.float64))
    a = init + leaky * init.sum() * init
    for t in range(T):
        M = torch.zeros(S, S, dtype=torch.float64)
        M = M.index_put((fsrc, fdst), fw * torch.exp(ll_t[t, fpdf]), accumulate=True)
        a = a @ M
        a = a + leaky * a.sum() * init
    return torch.log(a.sum())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_den_bruteforce_no_leaky(seed):
    S, N, T = 5, 4, 4
    g = chain_ref.den_graph_from_fst(tiny_den(S, N, seed), N)
    ll = np.random.default_rng(seed).normal(0, 1, (T, N))
    logZ_b, gam_b = brute_den(ll, g)
    logZ, gam, ok = chain_ref.den_fb_scaled(ll, g, leaky=0.0)
    assert ok
    np.testing.assert_allclose(logZ, logZ_b, rtol=1e-12)
    np.testing.assert_allclose(gam, gam_b, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("leaky", [0.0, 1e-4, 0.05])
def test_den_autograd_dense(leaky):
    S, N, T = 6, 5, 5
    g = chain_ref.den_graph_from_fst(tiny_den(S, N, 7), N)
    ll = np.random.default_rng(3).normal(0, 1.5, (T, N))
    llt = torch.tensor(ll, requires_grad=True)
    lz = dense_logZ(llt, g, leaky)
    lz.backward()
    logZ, gam, ok = chain_ref.den_fb_scaled(ll, g, leaky=leaky)
    np.testing.assert_allclose(logZ, lz.item(), rtol=1e-12)
    np.testing.assert_allclose(gam, llt.grad.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(gam.sum(1), 1.0, rtol=1e-10)


@pytest.mark.parametrize("leaky", [0.0, 1e-4])
def test_den_scaled_vs_log(leaky):
    N = 50
    fst = synth.make_den_fst(num_states=64, num_pdfs=N, mean_extra=5, seed=11)
    g = chain_ref.den_graph_from_fst(fst, N)
    ll = np.random.default_rng(5).normal(0, 2.0, (30, N))
    z1, g1, ok = chain_ref.den_fb_scaled(ll, g, leaky)
    z2, g2 = chain_ref.den_fb_log(ll, g, leaky)
    np.testing.assert_allclose(z1, z2, rtol=1e-10)
    np.testing.assert_allclose(g1, g2, rtol=1e-8, atol=1e-13)
    # fp32 variant of the same recursion stays within the parity budget
    z3, g3, _ = chain_ref.den_fb_scaled(ll, g, leaky, dtype=np.float32)
    np.testing.assert_allclose(z3, z1, rtol=1e-4)
    np.testing.assert_allclose(g3, g1, rtol=1e-3, atol=1e-6)


def test_initial_probs_and_csr():
    N = 20
    fst = synth.make_den_fst(num_states=32, num_pdfs=N, mean_extra=4, seed=2)
    g = chain_ref.den_graph_from_fst(fst, N)
    assert abs(g["initial_probs"].sum() - 1.0) < 1e-5
    # backward CSR is the transpose of the forward CSR
    S = g["num_states"]
    fsrc = np.repeat(np.arange(S), np.diff(g["fwd_off"]))
    bdst = np.repeat(np.arange(S), np.diff(g["bwd_off"]))
    f = sorted(zip(fsrc, g["fwd_state"], g["fwd_pdf"], g["fwd_prob"]))
    b = sorted(zip(g["bwd_state"], bdst, g["bwd_pdf"], g["bwd_prob"]))
    assert f == b


def brute_num(ll, fst):
    T, N = ll.shape
    out = [[] for _ in range(fst["num_states"])]
    for s, d, l, w in zip(fst["src"], fst["dst"], fst["ilabel"], fst["weight"]):
        out[s].append((d, l - 1, w))
    Z = 0.0
    num = np.zeros((T, N))

    def rec(s, t, p, pdfs):
        nonlocal Z
        if np.isfinite(fst["final"][s]):
            pp = p * np.exp(-fst["final"][s])
            Z += pp
            for tt, q in enumerate(pdfs):
                num[tt, q] += pp
        for (d, q, w) in out[s]:
            rec(d, t + 1, p * np.exp(ll[t, q] - w), pdfs + [q])

    rec(fst["start"], 0, 1.0, [])
    return np.log(Z), num / Z


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_num_bruteforce(seed):
    rng = np.random.default_rng(seed)
    T, N = 9, 6
    fst = synth.make_supervision_fst(T, N, rng, slack=1, min_dur=2, max_dur=3)
    times = chain_ref.fst_state_times(fst)
    assert (np.diff(times) >= 0).all()
    ll = rng.normal(0, 1, (T, N))
    zb, gb = brute_num(ll, fst)
    z, g = chain_ref.num_fb_log(ll, fst)
    np.testing.assert_allclose(z, zb, rtol=1e-12)
    np.testing.assert_allclose(g, gb, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(g.sum(1), 1.0, rtol=1e-10)


def test_chain_objf_invariants():
    rng = np.random.default_rng(0)
    N, T = 40, 25
    den = chain_ref.den_graph_from_fst(synth.make_den_fst(64, N, 5, seed=3), N)
    sup = synth.make_supervision_fst(T, N, rng)
    ll = rng.normal(0, 2, (T, N))
    objf, grad, gx = chain_ref.chain_objf_and_deriv(ll, den, sup, xent_regularize=0.0)
    np.testing.assert_allclose(grad.sum(1), 0.0, atol=1e-9)
    objf2, grad2, gx2 = chain_ref.chain_objf_and_deriv(ll, den, sup, xent_regularize=0.1)
    assert objf2 == objf
    np.testing.assert_allclose(grad2, grad + 0.1 * gx, rtol=1e-12)
