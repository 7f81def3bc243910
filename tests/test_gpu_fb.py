"""GPU parity tests (run on the B200 box): CUDA kernels behind the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star): losses / log-likelihoods within 1e-3 relative,
posteriors within 1e-3 relative (+ small absolute floor for ~0 entries); index tensors bit-exact.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


# ------------------------------------------------------------------------- fbank ----
def test_fbank_matches_reference_golden(dev):
    from pykaldi2_b200.data import fbank
    gold = np.load(os.path.join(G, "fbank_golden.npz"))
    ex = fbank.FbankExtractor()
    wavs = [gold["wav%d" % i] for i in range(5)]
    feats, foff, foff_d = ex(wavs)
    feats_h = feats.cpu().numpy()
    for i in range(5):
        ref = gold["fbank%d" % i]
        ours = feats_h[foff[i]:foff[i + 1]]
        assert ours.shape == ref.shape
        np.testing.assert_allclose(ours, ref, rtol=1e-3, atol=2e-3)
    mean = fbank.utterance_means(feats, foff_d, 5)
    src, utt, Tout, lens = fbank.padded_rows(foff)
    x = fbank.gather_norm(feats, src, utt, mean).view(5, Tout, 80).cpu().numpy()
    for i in range(5):
        T = gold["cmn%d" % i].shape[0]
        np.testing.assert_allclose(x[i, :T], gold["cmn%d" % i], rtol=1e-3, atol=2e-3)
        assert (x[i, T:] == 0).all()
    # chunks of utterance 4 with CMN, vs the reference's _utt2seg output
    csrc, cutt, cu, cs = fbank.chunk_rows(foff)
    ch = fbank.gather_norm(feats, csrc, cutt, mean).view(-1, 80, 80).cpu().numpy()
    sel = ch[cu == 4]
    assert sel.shape == gold["seg4"].shape
    np.testing.assert_allclose(sel, gold["seg4"], rtol=1e-3, atol=2e-3)
    # global MVN
    mm = torch.from_numpy(gold["mvn_mean"].reshape(-1)).to(dev)
    mi = torch.from_numpy((1.0 / gold["mvn_std"]).reshape(-1).astype(np.float32)).to(dev)
    T4 = gold["mvn4"].shape[0]
    s4 = np.arange(foff[4], foff[5]).astype(np.int32)
    y = fbank.gather_norm(feats, s4, np.full(T4, 4, np.int32), mean, (mm, mi)).cpu().numpy()
    np.testing.assert_allclose(y, gold["mvn4"], rtol=2e-3, atol=5e-3)


def test_fbank_vs_oracle_random_lengths(dev):
    from oracle import fbank_ref
    from pykaldi2_b200.data import fbank, mel
    rng = np.random.default_rng(7)
    wavs = [(0.05 * rng.standard_normal(n)).astype(np.float32) for n in (401, 402, 561, 8000, 31999, 48000)]
    feats, foff, _ = fbank.FbankExtractor()(wavs)
    f = feats.cpu().numpy()
    W = mel.mel80_window()
    for i, w in enumerate(wavs):
        ref = fbank_ref.logfbank(w, W)
        assert foff[i + 1] - foff[i] == ref.shape[0] == fbank.num_frames(len(w))
        np.testing.assert_allclose(f[foff[i]:foff[i + 1]], ref, rtol=1e-3, atol=2e-3)


def test_chain_subsample_rows_bit_exact():
    from oracle import fbank_ref
    from pykaldi2_b200.data import fbank
    foff = np.array([0, 10, 17, 40])
    for shift in (0, 1, 2):
        src, utt, Tout, lens = fbank.padded_rows(foff, factor=3, shift=shift)
        Tmax = 23
        idx = fbank_ref.chain_subsample_index(Tmax, shift, 3)
        assert Tout == len(idx)
        src = src.reshape(3, Tout)
        for b in range(3):
            exp = [foff[b] + p if p < lens[b] else -1 for p in idx]
            assert src[b].tolist() == exp


# -------------------------------------------------------------------- CE softmax ----
def test_ce_softmax_matches_torch(dev):
    from pykaldi2_b200 import _lib
    torch.manual_seed(0)
    R, N = 300, 5768
    logits = torch.randn(R, N, device=dev) * 3
    labels = torch.randint(0, N, (R,), device=dev)
    labels[::7] = -100
    loss = torch.empty(R, device=dev)
    grad = torch.empty_like(logits)
    _lib.check(_lib.lib().pk2_ce_softmax(_lib.ptr(logits), _lib.ptr(labels), R, N, 0.5, _lib.ptr(loss),
                                         _lib.ptr(grad), _lib.stream()), "ce")
    lg = logits.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lg, labels, ignore_index=-100, reduction="sum")
    (0.5 * ref).backward()
    np.testing.assert_allclose(loss.sum().item(), ref.item(), rtol=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), lg.grad.float().cpu().numpy(), rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------ LF-MMI chain ----
def _chain_case(S, N, Ts, seed, mean_extra=5):
    from pykaldi2_b200 import synth
    rng = np.random.default_rng(seed)
    fst = synth.make_den_fst(S, N, mean_extra, seed=seed)
    sups = [synth.make_supervision_fst(T, N, rng) for T in Ts]
    Tmax = max(Ts)
    pred = rng.normal(0, 2.0, (len(Ts), Tmax, N)).astype(np.float32)
    return fst, sups, pred


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("S,N,Ts", [(64, 50, [20, 13, 1]), (1000, 333, [37, 50])])
def test_chain_objf_and_deriv_vs_oracle(dev, S, N, Ts, cluster):
    from oracle import chain_ref
    from pykaldi2_b200 import graphs
    from pykaldi2_b200.ops import ops
    fst, sup_fsts, pred = _chain_case(S, N, Ts, seed=S + cluster)
    den = graphs.DenominatorGraph(fst, N)
    oden = chain_ref.den_graph_from_fst(fst, N)
    # index tensors bit-exact
    for k in ("fwd_off", "fwd_pdf", "fwd_state", "bwd_off", "bwd_pdf", "bwd_state"):
        assert (getattr(den, k) == oden[k]).all(), k
    assert (den.fwd_prob == oden["fwd_prob"]).all() and (den.initial_probs == oden["initial_probs"]).all()
    sups = [graphs.Supervision(f, T, N) for f, T in zip(sup_fsts, Ts)]
    sb = graphs.SupervisionBatch(sups, device=dev)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.1)
    p = torch.from_numpy(pred).to(dev)
    objf, grad = ops.chain_objf_and_deriv(p, den, sb, opts, cluster=cluster)
    objf = objf.cpu().numpy()
    grad = grad.cpu().numpy()
    for b, T in enumerate(Ts):
        o, g, gx = chain_ref.chain_objf_and_deriv(pred[b, :T], oden, sup_fsts[b], leaky=1e-4, xent_regularize=0.1)
        np.testing.assert_allclose(objf[b], o, rtol=1e-3)
        np.testing.assert_allclose(grad[b, :T], -g, rtol=1e-3, atol=2e-6)
        assert (grad[b, T:] == 0).all()


@pytest.mark.parametrize("S,N,Ts,extra", [(256, 52, [20, 13, 1, 7, 33], 5),          # 5 clusters, one sequence each
                                          (1024, 332, [37, 50], 14),                 # rows longer than the register slots
                                          (256, 76, list(range(1, 41)), 5)])         # work lists: more sequences than clusters
def test_chain_cluster8_register_resident(dev, S, N, Ts, extra):
    """Clusters of 8 CTAs with register-resident arcs (denfb.cu: den_*_reg_kernel) against the oracle."""
    from oracle import chain_ref
    from pykaldi2_b200 import graphs
    from pykaldi2_b200.ops import ops
    fst, sup_fsts, pred = _chain_case(S, N, Ts, seed=S + 8, mean_extra=extra)
    den = graphs.DenominatorGraph(fst, N)
    oden = chain_ref.den_graph_from_fst(fst, N)
    sups = [graphs.Supervision(f, T, N) for f, T in zip(sup_fsts, Ts)]
    sb = graphs.SupervisionBatch(sups, device=dev)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.1)
    p = torch.from_numpy(pred).to(dev)
    objf, grad = ops.chain_objf_and_deriv(p, den, sb, opts, cluster=8)
    objf = objf.cpu().numpy()
    grad = grad.cpu().numpy()
    for b, T in enumerate(Ts):
        o, g, gx = chain_ref.chain_objf_and_deriv(pred[b, :T], oden, sup_fsts[b], leaky=1e-4, xent_regularize=0.1)
        np.testing.assert_allclose(objf[b], o, rtol=1e-3)
        np.testing.assert_allclose(grad[b, :T], -g, rtol=1e-3, atol=2e-6)
        assert (grad[b, T:] == 0).all()
    # the automatic choice (cluster = 0): same kernels for the long sequences, the shortest ones of a large
    # batch as single-CTA kernels on the SMs the clusters leave free (hybrid schedule)
    objf0, grad0 = ops.chain_objf_and_deriv(p, den, sb, opts, cluster=0)
    np.testing.assert_allclose(objf0.cpu().numpy(), objf, rtol=1e-5)
    np.testing.assert_allclose(grad0.cpu().numpy(), grad, rtol=1e-3, atol=1e-6)


def test_chain_function_per_utt_and_batch(dev):
    from oracle import chain_ref
    from pykaldi2_b200 import graphs
    from pykaldi2_b200.ops import ops
    S, N, Ts = 128, 90, [25, 31]
    fst, sup_fsts, pred = _chain_case(S, N, Ts, seed=5)
    den = graphs.DenominatorGraph(fst, N)
    oden = chain_ref.den_graph_from_fst(fst, N)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
    sups = [graphs.Supervision(f, T, N) for f, T in zip(sup_fsts, Ts)]
    p = torch.from_numpy(pred).to(dev).requires_grad_(True)
    loss = 0.0
    for j, T in enumerate(Ts):                      # the reference's calling pattern
        loss = loss + ops.ChainObjtiveFunction.apply(p[j, :T, :], den, sups[j], opts)
    assert loss.device.type == "cpu" and loss.dim() == 0
    loss.backward()
    g1 = p.grad.clone()
    p.grad = None
    loss_b = ops.ChainObjtiveFunction.apply_batch(p, den, sups, opts)
    loss_b.backward()
    np.testing.assert_allclose(loss.item(), loss_b.item(), rtol=1e-6)
    np.testing.assert_allclose(g1.cpu().numpy(), p.grad.cpu().numpy(), rtol=1e-5, atol=1e-7)
    tot = 0.0
    for j, T in enumerate(Ts):
        o, g, _ = chain_ref.chain_objf_and_deriv(pred[j, :T], oden, sup_fsts[j], leaky=1e-4)
        tot += o
        np.testing.assert_allclose(g1[j, :T].cpu().numpy(), -g, rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(loss.item(), tot, rtol=1e-3)


def test_denfb_full_size_properties(dev):
    """BASELINE config 4 sizes (S=8192, N=5768): size-independent properties."""
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    N, S = 5768, 8192
    rng = np.random.default_rng(1234)
    den = graphs.DenominatorGraph(synth.make_den_fst(S, N, 7, seed=1234), N)
    Ts = [150, 97, 33, 150]
    sups = [graphs.Supervision(synth.make_supervision_fst(T, N, rng), T, N) for T in Ts]
    sb = graphs.SupervisionBatch(sups, device=dev)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
    pred = torch.from_numpy(rng.normal(0, 2.0, (len(Ts), max(Ts), N)).astype(np.float32)).to(dev)
    res = {}
    for K in (1, 2, 4, 8):
        objf, grad = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=K)
        res[K] = (objf.cpu().numpy(), grad.cpu().numpy())
    o1, g1 = res[1]
    assert np.isfinite(o1).all()
    for b, T in enumerate(Ts):
        rows = g1[b, :T].astype(np.float64).sum(1)      # gamma_den - gamma_num rows sum to 0
        np.testing.assert_allclose(rows, 0.0, atol=2e-4)
        assert (g1[b, T:] == 0).all()
    for K in (2, 4, 8):                                 # cluster size does not change the answer
        np.testing.assert_allclose(res[K][0], o1, rtol=1e-5)
        np.testing.assert_allclose(res[K][1], g1, rtol=1e-3, atol=1e-6)
    # shifting all loglikes of a frame by c shifts objf by 0 (num and den both move by c)
    pred2 = pred.clone()
    pred2[:, 5, :] += 1.5
    objf2, _ = ops.chain_objf_and_deriv(pred2, den, sb, opts, cluster=2)
    np.testing.assert_allclose(objf2.cpu().numpy(), o1, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("cluster", [0, 8, 4, 1])
def test_chain_full_size_vs_oracle(dev, cluster):
    """BASELINE config 4 sizes (S = 8192, N = 5768, mean out-degree 8) against the fp64 oracle: objective and
    derivatives of every sequence.  26 sequences of 30..150 frames: with cluster = 0 the automatic schedule puts the
    long ones on clusters of 8 (register-resident arcs, two rows per thread) and the short ones on single-CTA
    kernels; 8 / 4 / 1 force the register-resident, the 4-CTA streaming and the 1-CTA streaming kernels."""
    from oracle import chain_ref
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    N, S = 5768, 8192
    rng = np.random.default_rng(4321)
    fst = synth.make_den_fst(S, N, 7, seed=1234)
    den = graphs.DenominatorGraph(fst, N)
    oden = chain_ref.den_graph_from_fst(fst, N)
    for k in ("fwd_off", "fwd_pdf", "fwd_state", "bwd_off", "bwd_pdf", "bwd_state"):
        assert (getattr(den, k) == oden[k]).all(), k
    assert (den.fwd_prob == oden["fwd_prob"]).all() and (den.initial_probs == oden["initial_probs"]).all()
    Ts = [150, 30] + [int(t) for t in rng.integers(30, 151, size=24)]
    sup_fsts = [synth.make_supervision_fst(T, N, rng) for T in Ts]
    sups = [graphs.Supervision(f, T, N) for f, T in zip(sup_fsts, Ts)]
    sb = graphs.SupervisionBatch(sups, device=dev)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
    pred = rng.normal(0, 2.0, (len(Ts), max(Ts), N)).astype(np.float32)
    objf, grad = ops.chain_objf_and_deriv(torch.from_numpy(pred).to(dev), den, sb, opts, cluster=cluster)
    objf, grad = objf.cpu().numpy(), grad.cpu().numpy()
    for b, T in enumerate(Ts):
        o, g, _ = chain_ref.chain_objf_and_deriv(pred[b, :T], oden, sup_fsts[b], leaky=1e-4)
        np.testing.assert_allclose(objf[b], o, rtol=1e-3, err_msg="sequence %d (T=%d)" % (b, T))
        np.testing.assert_allclose(grad[b, :T], -g, rtol=1e-3, atol=2e-6, err_msg="sequence %d (T=%d)" % (b, T))
        assert (grad[b, T:] == 0).all()


def test_denfb_full_size_hybrid_schedule_matches_streaming(dev):
    """BASELINE config 4 graph, more sequences than resident clusters: the automatic schedule (clusters of 8 with
    work lists + single-CTA kernels on the spare SMs) against the streaming 4-CTA kernels."""
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    N, S = 5768, 8192
    rng = np.random.default_rng(99)
    den = graphs.DenominatorGraph(synth.make_den_fst(S, N, 7, seed=1234), N)
    Ts = [int(t) for t in rng.integers(8, 60, size=26)]
    sups = [graphs.Supervision(synth.make_supervision_fst(T, N, rng), T, N) for T in Ts]
    sb = graphs.SupervisionBatch(sups, device=dev)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
    pred = torch.from_numpy(rng.normal(0, 2.0, (len(Ts), max(Ts), N)).astype(np.float32)).to(dev)
    o0, g0 = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=0)
    o4, g4 = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=4)
    np.testing.assert_allclose(o0.cpu().numpy(), o4.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(g0.cpu().numpy(), g4.cpu().numpy(), rtol=1e-3, atol=1e-6)
    for b, T in enumerate(Ts):
        assert (g0[b, T:] == 0).all()


def test_denfb_random_batches_two_graphs_alternating(dev):
    """Random batch sizes / lengths on two graphs of different size used alternately (the kernels' shared-memory
    attribute is per function, not per graph): automatic and cluster-8 schedules against the streaming kernels."""
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(21)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
    dens = [(graphs.DenominatorGraph(synth.make_den_fst(S, N, 7, seed=S), N), N) for S, N in ((2048, 400), (256, 52))]
    for trial in range(6):
        den, N = dens[trial % 2]
        B = int(rng.integers(1, 60))
        Ts = [int(t) for t in rng.integers(1, 70, size=B)]
        sups = [graphs.Supervision(synth.make_supervision_fst(T, N, rng), T, N) for T in Ts]
        sb = graphs.SupervisionBatch(sups, device=dev)
        pred = torch.from_numpy(rng.normal(0, 2.0, (B, max(Ts), N)).astype(np.float32)).to(dev)
        o1, g1 = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=[1, 2, 4][trial % 3])
        for K in (0, 8):
            o, g = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=K)
            np.testing.assert_allclose(o.cpu().numpy(), o1.cpu().numpy(), rtol=1e-5)
            np.testing.assert_allclose(g.cpu().numpy(), g1.cpu().numpy(), rtol=1e-3, atol=1e-6)


# ------------------------------------------------------------------- lattice MMI ----
@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_lattice_mmi_vs_oracle(dev, eps):
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(11)
    N, Ts = 200, [40, 23, 31]
    lats, alis, olat = [], [], []
    for T in Ts:
        lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=8, kmax=20, ali_drop=0.2, eps_frac=eps)
        olat.append(lat); alis.append(ali); lats.append(graphs.Lattice(lat))
    pred = rng.normal(0, 3.0, (len(Ts), max(Ts), N)).astype(np.float32)
    lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev)
    tot, grad = ops.lattice_mmi(torch.from_numpy(pred).to(dev), lb)
    tot, grad = tot.cpu().numpy(), grad.cpu().numpy()
    for b, T in enumerate(Ts):
        rtot, post, drop, times = lattice_ref.lattice_fb_mmi(pred[b, :T], olat[b], tid2pdf, alis[b])
        assert (lats[b].state_times_orig == times).all()          # index tensors bit-exact
        assert (lb.keep_host[b] == (~drop).astype(np.uint8)).all()
        np.testing.assert_allclose(tot[b], rtot, rtol=1e-6)
        np.testing.assert_allclose(grad[b, :T], -post, rtol=1e-3, atol=1e-6)
        assert (grad[b, T:] == 0).all()


def test_mmi_function_reference_call_pattern(dev):
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(3)
    N, T = 120, 35
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=6, kmax=12)
    trans_model = graphs.TidPdfMap(tid2pdf)
    decoder = graphs.SyntheticLatticeProvider([graphs.Lattice(lat)])
    log_prior = torch.from_numpy(synth.make_log_prior(N, rng)).to(dev)
    logits = torch.randn(T, N, device=dev, requires_grad=True)
    loglike = logits - log_prior                       # bin/train_se.py:241
    loss = ops.MMIFunction.apply(loglike, decoder, trans_model, ali.tolist())
    assert loss.device.type == "cpu"
    loss.backward()
    ll = (logits.detach() - log_prior).cpu().numpy()
    rtot, post, drop, _ = lattice_ref.lattice_fb_mmi(ll, lat, tid2pdf, ali)
    np.testing.assert_allclose(loss.item(), rtot, rtol=1e-5)
    np.testing.assert_allclose(logits.grad.cpu().numpy(), -post, rtol=1e-3, atol=1e-6)


# ------------------------------------------------------------------- sMBR / MPFE ----
@pytest.mark.parametrize("criterion", ["smbr", "mpfe"])
@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_lattice_mpe_vs_oracle(dev, eps, criterion):
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(17)
    N, Ts = 200, [40, 23, 31]
    lats, alis, olat = [], [], []
    for T in Ts:
        lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=8, kmax=20, ali_drop=0.2, eps_frac=eps)
        olat.append(lat); alis.append(ali); lats.append(graphs.Lattice(lat))
    tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 3 + 1, 0)
    sil = [1, 2]
    pred = rng.normal(0, 1.0, (len(Ts), max(Ts), N)).astype(np.float32)
    lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev, mpe=(criterion, tid2phone, sil))
    score, grad, tot = ops.lattice_mpe(torch.from_numpy(pred).to(dev), lb)
    score, grad, tot = score.cpu().numpy(), grad.cpu().numpy(), tot.cpu().numpy()
    for b, T in enumerate(Ts):
        rs, post, rtot = lattice_ref.lattice_fb_mpe(pred[b, :T], olat[b], tid2pdf, tid2phone, alis[b], criterion, sil)
        # per-arc frame accuracies are index work: bit-exact against the oracle's per-arc function
        L = lats[b]
        t_out = np.repeat(L.state_time[:L.num_states], np.diff(L.out_off))
        ref = [lattice_ref.mpe_frame_acc(int(t), int(alis[b][tt]), tid2pdf, tid2phone, criterion, set(sil))
               for t, tt in zip(L.out_tid, t_out)]
        assert (np.asarray(ref, np.uint8) == L.frame_acc(alis[b], tid2pdf, tid2phone, criterion, sil)[1]).all()
        np.testing.assert_allclose(tot[b], rtot, rtol=1e-6)
        np.testing.assert_allclose(score[b], rs, rtol=1e-5)
        np.testing.assert_allclose(grad[b, :T], -post, rtol=1e-3, atol=1e-6)
        assert (grad[b, T:] == 0).all()


def test_smbr_function_reference_call_pattern(dev):
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(5)
    N, T = 120, 35
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=6, kmax=12)
    tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 3 + 1, 0)
    trans_model = graphs.TidPdfMap(tid2pdf, tid2phone)
    decoder = graphs.SyntheticLatticeProvider([graphs.Lattice(lat)])
    log_prior = torch.from_numpy(synth.make_log_prior(N, rng)).to(dev)
    logits = torch.randn(T, N, device=dev, requires_grad=True)
    loglike = logits - log_prior                       # bin/train_se.py:241
    loss = ops.sMBRFunction.apply(loglike, decoder, trans_model, ali.tolist(), "smbr", [1])   # bin/train_se.py:249
    assert loss.device.type == "cpu" and loss.dim() == 0
    loss.backward()
    ll = (logits.detach() - log_prior).cpu().numpy()
    rs, post, _ = lattice_ref.lattice_fb_mpe(ll, lat, tid2pdf, tid2phone, ali, "smbr", [1])
    np.testing.assert_allclose(loss.item(), rs, rtol=1e-5)
    np.testing.assert_allclose(logits.grad.cpu().numpy(), -post, rtol=1e-3, atol=1e-6)
    # batched entry used by bin/train_se.py -batched_loss 1
    lb = graphs.LatticeBatch([graphs.Lattice(lat)], tid2pdf, [ali], device=dev, mpe=("smbr", tid2phone, [1]))
    logits2 = logits.detach().clone().requires_grad_(True)
    loss2 = ops.sMBRFunction.apply_batch((logits2 - log_prior).unsqueeze(0), lb)
    loss2.backward()
    np.testing.assert_allclose(loss2.item(), loss.item(), rtol=1e-6)
    np.testing.assert_allclose(logits2.grad.cpu().numpy(), logits.grad.cpu().numpy(), rtol=1e-5, atol=1e-7)


# ------------------------------------------------- lattice kernels at BASELINE config 3 scale ----
def _c3_lattices(rng, N, Ts, eps):
    from pykaldi2_b200 import graphs, synth
    lats, alis, olat = [], [], []
    for T in Ts:
        lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=32, kmax=96, ali_drop=0.05, eps_frac=eps)
        olat.append(lat); alis.append(ali); lats.append(graphs.Lattice(lat))
    return lats, alis, olat, tid2pdf


@pytest.mark.parametrize("eps", [0.0, 0.05])
def test_lattice_mmi_c3_scale_vs_oracle(dev, eps):
    """SURVEY 8d C3 shape: N = 5768, T >= 600, K_t ~ U{32..96} states per frame (about 250 arcs per frame), 5 % of
    the frames without the alignment arc (drop_frames), optional 5 % epsilon arcs."""
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(31)
    N, Ts = 5768, [640, 905]
    lats, alis, olat, tid2pdf = _c3_lattices(rng, N, Ts, eps)
    pred = rng.normal(0, 3.0, (len(Ts), max(Ts), N)).astype(np.float32)
    lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev)
    tot, grad = ops.lattice_mmi(torch.from_numpy(pred).to(dev), lb)
    tot, grad = tot.cpu().numpy(), grad.cpu().numpy()
    for b, T in enumerate(Ts):
        rtot, post, drop, times = lattice_ref.lattice_fb_mmi(pred[b, :T], olat[b], tid2pdf, alis[b])
        assert (lats[b].state_times_orig == times).all()
        assert (lb.keep_host[b] == (~drop).astype(np.uint8)).all()
        assert drop.any() and not drop.all()
        np.testing.assert_allclose(tot[b], rtot, rtol=1e-6)
        np.testing.assert_allclose(grad[b, :T], -post, rtol=1e-3, atol=1e-6)
        assert (grad[b, T:] == 0).all()


@pytest.mark.parametrize("criterion,eps", [("smbr", 0.0), ("smbr", 0.05), ("mpfe", 0.05)])
def test_lattice_mpe_c3_scale_vs_oracle(dev, criterion, eps):
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(37)
    N, Ts = 5768, [600, 811]
    lats, alis, olat, tid2pdf = _c3_lattices(rng, N, Ts, eps)
    tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 40 + 1, 0)
    sil = [1, 2]
    pred = rng.normal(0, 1.0, (len(Ts), max(Ts), N)).astype(np.float32)
    lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev, mpe=(criterion, tid2phone, sil))
    score, grad, tot = ops.lattice_mpe(torch.from_numpy(pred).to(dev), lb)
    score, grad, tot = score.cpu().numpy(), grad.cpu().numpy(), tot.cpu().numpy()
    for b, T in enumerate(Ts):
        rs, post, rtot = lattice_ref.lattice_fb_mpe(pred[b, :T], olat[b], tid2pdf, tid2phone, alis[b], criterion, sil)
        np.testing.assert_allclose(tot[b], rtot, rtol=1e-6)
        np.testing.assert_allclose(score[b], rs, rtol=1e-5)
        np.testing.assert_allclose(grad[b, :T], -post, rtol=1e-3, atol=1e-6)
        assert (grad[b, T:] == 0).all()


@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_lattice_wide_levels_vs_oracle(dev, eps):
    """Levels wider than a chain CTA (more than 128 states per frame) and states with more arcs than the register
    prefetch holds: the kernels' memory path, MMI and sMBR."""
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(41)
    N, Ts = 300, [23, 9, 1]
    lats, alis, olat = [], [], []
    for T in Ts:
        lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=140, kmax=300, ali_drop=0.2, eps_frac=eps)
        olat.append(lat); alis.append(ali); lats.append(graphs.Lattice(lat))
    assert max(int(np.diff(l.level_off).max()) for l in lats) > 128
    assert max(int(np.diff(l.in_off).max()) for l in lats) > 6
    tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 3 + 1, 0)
    pred = rng.normal(0, 2.0, (len(Ts), max(Ts), N)).astype(np.float32)
    p = torch.from_numpy(pred).to(dev)
    lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev, mpe=("smbr", tid2phone, [1, 2]))
    tot, grad = ops.lattice_mmi(p, lb)
    score, grad_s, tot_s = ops.lattice_mpe(p, lb)
    tot, grad, score, grad_s, tot_s = (x.cpu().numpy() for x in (tot, grad, score, grad_s, tot_s))
    for b, T in enumerate(Ts):
        rtot, post, drop, _ = lattice_ref.lattice_fb_mmi(pred[b, :T], olat[b], tid2pdf, alis[b])
        np.testing.assert_allclose(tot[b], rtot, rtol=1e-6)
        np.testing.assert_allclose(grad[b, :T], -post, rtol=1e-3, atol=1e-6)
        rs, post_s, rtot_s = lattice_ref.lattice_fb_mpe(pred[b, :T], olat[b], tid2pdf, tid2phone, alis[b], "smbr", [1, 2])
        np.testing.assert_allclose(tot_s[b], rtot_s, rtol=1e-6)
        np.testing.assert_allclose(score[b], rs, rtol=1e-5)
        np.testing.assert_allclose(grad_s[b, :T], -post_s, rtol=1e-3, atol=1e-6)


def test_lattice_batch_validates_indices(dev):
    """Transition ids outside the tid -> pdf map and pdfs outside the prediction's columns are refused on the host
    (the kernels index with them)."""
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    rng = np.random.default_rng(2)
    N, T = 50, 12
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=4, kmax=8)
    with pytest.raises(ValueError):
        graphs.LatticeBatch([graphs.Lattice(lat)], tid2pdf[:20], [ali], device=dev)
    lb = graphs.LatticeBatch([graphs.Lattice(lat)], tid2pdf, [ali], device=dev)
    with pytest.raises(RuntimeError):
        ops.lattice_mmi(torch.zeros(1, T, N - 10, device=dev), lb)
