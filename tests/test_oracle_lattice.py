"""Pin the lattice-MMI oracle (oracle/lattice_ref.py) by brute-force path
enumeration, autograd, and hand-built merge cases (cancel / drop_frames / epsilon)."""
import numpy as np
import pytest
import torch

from oracle import lattice_ref
from pykaldi2_b200 import synth


def brute(ll, lat, tid2pdf, lm=1.0, ac=0.2):
    T, N = ll.shape
    S = lat["num_states"]
    out = [[] for _ in range(S)]
    for s, d, l, g in zip(lat["src"], lat["dst"], lat["tid"], lat["graph_cost"]):
        out[s].append((int(d), int(l), float(g)))
    Z = 0.0
    post = [dict() for _ in range(T)]

    def rec(s, t, logp, lab):
        nonlocal Z
        if np.isfinite(lat["final_cost"][s]) and t == T:
            p = np.exp(logp - lm * lat["final_cost"][s])
            Z += p
            for tt, l in enumerate(lab):
                post[tt][l] = post[tt].get(l, 0.0) + p
        for (d, l, g) in out[s]:
            if l == 0:
                rec(d, t, logp - lm * g, lab)
            elif t < T:
                rec(d, t + 1, logp - lm * g + ac * ll[t, tid2pdf[l]], lab + [l])

    rec(0, 0, 0.0, [])
    for t in range(T):
        for k in post[t]:
            post[t][k] /= Z
    return np.log(Z), post


@pytest.mark.parametrize("seed,eps", [(0, 0.0), (1, 0.0), (2, 0.3), (3, 0.3)])
def test_bruteforce(seed, eps):
    rng = np.random.default_rng(seed)
    T, N = 4, 5
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=2, kmax=3, dmin=1, dmax=3,
                                           ali_drop=0.3, eps_frac=eps)
    ll = rng.normal(0, 2, (T, N))
    tot, post, drop, times = lattice_ref.lattice_fb_mmi(ll, lat, tid2pdf, ali)
    zb, pb = brute(ll, lat, tid2pdf)
    np.testing.assert_allclose(tot, zb, rtol=1e-12)
    exp = np.zeros((T, N))
    for t in range(T):
        if int(ali[t]) not in pb[t]:
            assert drop[t]
            continue
        assert not drop[t]
        exp[t, tid2pdf[ali[t]]] += 1.0
        for k, v in pb[t].items():
            exp[t, tid2pdf[k]] -= v
    np.testing.assert_allclose(post, exp, rtol=1e-9, atol=1e-13)


def test_autograd_matches_den_posterior():
    """den posterior * ac_scale == d tot / d loglikes."""
    rng = np.random.default_rng(5)
    T, N = 6, 7
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=3, kmax=5, dmin=2, dmax=3, ali_drop=0.0)
    ll = rng.normal(0, 1, (T, N))
    times = lattice_ref.lattice_state_times(lat)
    llt = torch.tensor(ll, requires_grad=True)
    S = lat["num_states"]
    alpha = [torch.tensor(-np.inf, dtype=torch.float64)] * S
    alpha[0] = torch.tensor(0.0, dtype=torch.float64)
    for s, d, l, g in zip(lat["src"], lat["dst"], lat["tid"], lat["graph_cost"]):
        like = -float(g) + (0.2 * llt[times[s], tid2pdf[l]] if l != 0 else 0.0)
        alpha[d] = torch.logaddexp(alpha[d], alpha[s] + like)
    fin = [alpha[s] - float(lat["final_cost"][s]) for s in range(S) if np.isfinite(lat["final_cost"][s])]
    tot_t = torch.logsumexp(torch.stack(fin), 0)
    tot_t.backward()
    tot, post, drop, _ = lattice_ref.lattice_fb_mmi(ll, lat, tid2pdf, ali)
    np.testing.assert_allclose(tot, tot_t.item(), rtol=1e-12)
    assert not drop.any()
    num = np.zeros((T, N))
    num[np.arange(T), tid2pdf[ali]] = 1.0
    den = num - post
    np.testing.assert_allclose(den * 0.2, llt.grad.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(den.sum(1), 1.0, rtol=1e-10)


def test_merge_semantics_handbuilt():
    # one frame, two parallel arcs: tid 1 and tid 3 (pdfs 0 and 1)
    lat = {"num_states": 2, "src": np.array([0, 0]), "dst": np.array([1, 1]),
           "tid": np.array([1, 3]), "graph_cost": np.array([0.0, 0.0], np.float32),
           "final_cost": np.array([np.inf, 0.0], np.float32)}
    tid2pdf = np.array([-1, 0, 0, 1, 1])
    ll = np.zeros((1, 2))
    # alignment on tid 1: post = 1 - 0.5 on pdf 0, -0.5 on pdf 1
    tot, post, drop, times = lattice_ref.lattice_fb_mmi(ll, lat, tid2pdf, [1])
    np.testing.assert_allclose(post, [[0.5, -0.5]])
    assert times.tolist() == [0, 1] and not drop[0]
    # alignment tid 2 maps to the same pdf 0 but is a different tid -> disjoint -> frame dropped
    tot, post, drop, _ = lattice_ref.lattice_fb_mmi(ll, lat, tid2pdf, [2])
    assert drop[0] and (post == 0).all()
    # single-arc lattice equal to the alignment: exact cancellation -> zero row, not dropped
    lat1 = {"num_states": 2, "src": np.array([0]), "dst": np.array([1]), "tid": np.array([1]),
            "graph_cost": np.array([0.3], np.float32), "final_cost": np.array([np.inf, 0.1], np.float32)}
    tot, post, drop, _ = lattice_ref.lattice_fb_mmi(ll, lat1, tid2pdf, [1])
    assert not drop[0] and (post == 0).all()
    np.testing.assert_allclose(tot, -0.3 - 0.1, rtol=1e-6)


# ------------------------------------------------------------------ sMBR / MPFE ----
def brute_mpe(ll, lat, tid2pdf, tid2phone, ali, criterion, sil, lm=1.0, ac=1.0):
    """All paths: expected frame accuracy and sum_paths P(path) * (acc(path) - E[acc]) per (t, pdf)."""
    T, N = ll.shape
    S = lat["num_states"]
    out = [[] for _ in range(S)]
    for s, d, l, g in zip(lat["src"], lat["dst"], lat["tid"], lat["graph_cost"]):
        out[s].append((int(d), int(l), float(g)))
    paths = []

    def rec(s, t, logp, lab):
        if np.isfinite(lat["final_cost"][s]) and t == T:
            paths.append((np.exp(logp - lm * lat["final_cost"][s]), list(lab)))
        for (d, l, g) in out[s]:
            if l == 0:
                rec(d, t, logp - lm * g, lab)
            elif t < T:
                rec(d, t + 1, logp - lm * g + ac * ll[t, tid2pdf[l]], lab + [l])

    rec(0, 0, 0.0, [])
    Z = sum(p for p, _ in paths)
    accs = [sum(lattice_ref.mpe_frame_acc(l, int(ali[t]), tid2pdf, tid2phone, criterion, sil) for t, l in enumerate(lab))
            for _, lab in paths]
    E = sum(p * a for (p, _), a in zip(paths, accs)) / Z
    post = np.zeros((T, N))
    for (p, lab), a in zip(paths, accs):
        for t, l in enumerate(lab):
            post[t, tid2pdf[l]] += p / Z * (a - E)
    return E, post


@pytest.mark.parametrize("criterion", ["smbr", "mpfe"])
@pytest.mark.parametrize("seed,eps", [(0, 0.0), (1, 0.0), (2, 0.3), (3, 0.3)])
def test_mpe_bruteforce(seed, eps, criterion):
    rng = np.random.default_rng(100 + seed)
    T, N = 4, 6
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=2, kmax=3, dmin=1, dmax=3, ali_drop=0.3, eps_frac=eps)
    tid2phone = np.asarray(tid2pdf) // 2 + 1            # 3 phones, phone 1 = silence
    ll = rng.normal(0, 1, (T, N))
    score, post, tot = lattice_ref.lattice_fb_mpe(ll, lat, tid2pdf, tid2phone, ali, criterion, [1])
    E, pb = brute_mpe(ll, lat, tid2pdf, tid2phone, ali, criterion, {1})
    np.testing.assert_allclose(score, E, rtol=1e-10)
    np.testing.assert_allclose(post, pb, rtol=1e-8, atol=1e-12)
    assert 0.0 <= score <= T


def test_mpe_posterior_is_the_derivative_of_the_expected_accuracy():
    """d E[acc] / d loglike[t, p] = acoustic_scale * post[t, p] (central differences)."""
    rng = np.random.default_rng(7)
    T, N = 5, 6
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=2, kmax=4, dmin=1, dmax=3)
    tid2phone = np.asarray(tid2pdf) // 2 + 1
    ll = rng.normal(0, 1, (T, N))
    ac = 0.7
    f = lambda x: lattice_ref.lattice_fb_mpe(x, lat, tid2pdf, tid2phone, ali, "smbr", [1], ac_scale=ac)[0]
    _, post, _ = lattice_ref.lattice_fb_mpe(ll, lat, tid2pdf, tid2phone, ali, "smbr", [1], ac_scale=ac)
    h = 1e-5
    for t in range(T):
        for p in range(N):
            d = np.zeros_like(ll); d[t, p] = h
            np.testing.assert_allclose((f(ll + d) - f(ll - d)) / (2 * h), ac * post[t, p], rtol=1e-5, atol=1e-8)


def test_mpe_silence_classes():
    ph = np.array([0, 1, 1, 2, 3])            # tid -> phone; phones 1, 2 are silence
    pdf = np.array([0, 0, 1, 2, 2])
    acc = lattice_ref.mpe_frame_acc
    assert acc(1, 2, pdf, ph, "smbr", {1, 2}) == 1.0            # different pdf, both silence
    assert acc(1, 2, pdf, ph, "smbr", set()) == 0.0
    assert acc(3, 4, pdf, ph, "smbr", {1}) == 1.0               # same pdf
    assert acc(3, 4, pdf, ph, "mpfe", {1}) == 0.0               # different phones
    assert acc(1, 3, pdf, ph, "mpfe", {1, 2}) == 1.0            # both silence phones
    assert acc(1, 1, pdf, ph, "smbr", {1}, one_silence_class=False) == 0.0   # old behaviour: silence never counts
