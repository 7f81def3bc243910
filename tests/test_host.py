"""CPU-side tests: the C-ABI library loads and exports every symbol include/pk2.h declares, the
host-side index logic (graphs, row maps, collate) is bit-exact against the oracle / the reference's
golden vectors, and the N>1 data-parallel path works on gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    from pykaldi2_b200 import _lib, build
    build.build(verbose=False)
    decl = set(re.findall(r"\b(pk2_[a-z0-9_]+)\s*\(", open(os.path.join(ROOT, "include", "pk2.h")).read()))
    assert decl == set(_lib.EXPORTS)
    L = _lib.lib()                       # dlopen + getattr of every symbol
    for name in decl:
        assert hasattr(L, name), name
    assert L.pk2_version() >= 100
    assert L.pk2_launch_count() == 0     # nothing launched: no GPU here


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pykaldi2_b200/ or bin/ may import it."""
    bad = []
    for base in ("pykaldi2_b200", "bin"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith(".py"):
                    src = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_cpu_tensors_are_rejected():
    from pykaldi2_b200.models.lstm import LSTMAM
    m = LSTMAM(80, 16, 64, 1, 0.0, True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 4, 80))
    assert "lstm.weight_hh_l0_reverse" in m.state_dict() and "output_layer.bias" in m.state_dict()


def test_lattice_prep_index_tensors_bit_exact():
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    rng = np.random.default_rng(11)
    for eps in (0.0, 0.15):
        lat, t2p, ali = synth.make_lattice(30, 100, rng, kmin=5, kmax=12, ali_drop=0.3, eps_frac=eps)
        L = graphs.Lattice(lat)
        assert (L.state_times_orig == lattice_ref.lattice_state_times(lat)).all()
        ll = rng.normal(0, 2, (30, 100))
        _, _, drop, _ = lattice_ref.lattice_fb_mmi(ll, lat, t2p, ali)
        assert (L.keep_mask(ali) == (~drop)).all()
        assert L.level_off[-1] == L.num_states and (np.diff(L.state_time) >= 0).all()
        assert len(L.out_dst) + len(L.eps_src) == len(lat["src"])


def test_mpe_frame_accuracies_bit_exact():
    """Host index work of sMBR / MPFE (graphs.Lattice.frame_acc) against the oracle's per-arc rule."""
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    rng = np.random.default_rng(13)
    lat, t2p, ali = synth.make_lattice(25, 60, rng, kmin=4, kmax=9, ali_drop=0.3, eps_frac=0.1)
    t2ph = np.where(np.asarray(t2p) >= 0, np.asarray(t2p) // 3 + 1, 0)
    L = graphs.Lattice(lat)
    S = L.num_states
    t_out = np.repeat(L.state_time[:S], np.diff(L.out_off))
    t_in = np.repeat(L.state_time[:S], np.diff(L.in_off)) - 1
    for crit in ("smbr", "mpfe"):
        for sil in ([], [1, 3]):
            a_in, a_out = L.frame_acc(ali, t2p, t2ph, crit, sil)
            r_out = [lattice_ref.mpe_frame_acc(int(t), int(ali[tt]), t2p, t2ph, crit, set(sil)) for t, tt in zip(L.out_tid, t_out)]
            r_in = [lattice_ref.mpe_frame_acc(int(t), int(ali[tt]), t2p, t2ph, crit, set(sil)) for t, tt in zip(L.in_tid, t_in)]
            assert (a_out == np.asarray(r_out, np.uint8)).all() and (a_in == np.asarray(r_in, np.uint8)).all()
    with pytest.raises(ValueError):
        L.frame_acc(ali, t2p, t2ph, "mmi", [])
    tm = graphs.TidPdfMap(t2p, t2ph)
    assert tm.transition_id_to_phone(5) == int(t2ph[5])
    with pytest.raises(RuntimeError):
        graphs.TidPdfMap(t2p).transition_id_to_phone(1)


def test_supervision_prep_and_generic_time_sort():
    from oracle import chain_ref
    from pykaldi2_b200 import graphs, synth
    rng = np.random.default_rng(2)
    f = synth.make_supervision_fst(25, 40, rng)
    s = graphs.Supervision(f, 25, 40)
    assert (s.state_time == chain_ref.fst_state_times(f)).all()
    g = dict(f)
    del g["state_times"]                 # generic path: times recomputed by BFS
    s2 = graphs.Supervision(g, 25, 40)
    for k in ("state_time", "out_off", "out_dst", "out_pdf", "in_off", "in_src", "in_pdf", "level_off"):
        assert (getattr(s, k) == getattr(s2, k)).all(), k
    with pytest.raises(ValueError):
        graphs.Supervision(f, 24, 40)


def test_row_maps_match_reference_semantics():
    from oracle import fbank_ref
    from pykaldi2_b200.data import fbank
    gold = np.load(os.path.join(G, "fbank_golden.npz"))
    for i in range(5):
        assert fbank.num_frames(len(gold["wav%d" % i])) == gold["fbank%d" % i].shape[0]
    foff = np.array([0, 250, 330, 331])
    src, utt, cu, cs = fbank.chunk_rows(foff)
    assert cu.tolist() == [0, 0, 0, 1] and cs.tolist() == [0, 80, 160, 0]
    assert cs.tolist()[:3] == fbank_ref.utt2seg_index(250)
    src, utt, Tout, lens = fbank.padded_rows(foff)
    assert Tout == 250 and src.reshape(3, 250)[2].tolist() == [330] + [-1] * 249


def test_collate_golden():
    from pykaldi2_b200.data import dataloader
    g = np.load(os.path.join(G, "collate_golden.npz"))
    items = [(g["feat%d" % j], ["utt%d" % j], g["lab%d" % j], [g["aux%d" % j]]) for j in range(3)]
    b = dataloader.seq_collate(items)
    assert (b["x"].numpy() == g["x"]).all() and (b["y"].numpy() == g["y"]).all()
    assert list(b["num_frs"]) == g["num_frs"].tolist()
    assert b["aux"][1][0][0].tolist() == g["aux1"][0].tolist()


def test_cmn_and_mvn_pickle_compat():
    import pickle
    from pykaldi2_b200.reader import preprocess
    g = np.load(os.path.join(G, "fbank_golden.npz"))
    np.testing.assert_allclose(preprocess.cmn(g["fbank2"], axis=0), g["cmn2"], rtol=1e-5, atol=1e-5)
    tr = preprocess.GlobalMeanVarianceNormalization()
    tr.mean_vec, tr.std_vec = g["mvn_mean"], g["mvn_std"]
    tr2 = pickle.loads(pickle.dumps(tr))
    np.testing.assert_allclose(tr2.apply_on_ndarray(g["cmn4"]), g["mvn4"], rtol=1e-6)


_DP_SCRIPT = r"""
import os, sys, torch
sys.path.insert(0, %r)
from pykaldi2_b200 import dist as pkdist
rank, world, local = pkdist.init(backend="gloo")
torch.manual_seed(0)
model = torch.nn.Linear(8, 4)
pkdist.broadcast_parameters(model)
opt = pkdist.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1))
x = torch.arange(16, dtype=torch.float32).view(2, 8) + 100 * rank
loss = model(x).pow(2).sum()
opt.zero_grad(); loss.backward(); opt.synchronize()
g = model.weight.grad.clone()
# reference: mean of the two ranks' gradients computed locally
ref = torch.zeros_like(g)
for r in range(world):
    m2 = torch.nn.Linear(8, 4); m2.load_state_dict(model.state_dict())
    xr = torch.arange(16, dtype=torch.float32).view(2, 8) + 100 * r
    m2(xr).pow(2).sum().backward(); ref += m2.weight.grad / world
assert torch.allclose(g, ref, rtol=1e-5, atol=1e-5), (g, ref)
opt.step()
w = model.weight.detach().clone()
ws = [torch.zeros_like(w) for _ in range(world)]
torch.distributed.all_gather(ws, w)
assert all(torch.equal(ws[0], t) for t in ws)
assert pkdist.shard_indices(5, world, rank) == [(rank + i * world) %% 5 for i in range(3)]
# GradAverager hands the averaged gradients back as views of its flat bucket (no copy back): distinct storage
# offsets per parameter, values = mean over ranks, clipping + a second step still work
m3 = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.Linear(4, 3))
pkdist.broadcast_parameters(m3)
avg = pkdist.GradAverager(list(m3.parameters()))
for it in range(2):
    m3(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in m3.parameters()]
    avg.average()
    for p, g in zip(m3.parameters(), local):
        both = [torch.zeros_like(g) for _ in range(world)]
        torch.distributed.all_gather(both, g)
        assert torch.allclose(p.grad, sum(both) / world, rtol=1e-6, atol=1e-6)
        assert p.grad.untyped_storage().data_ptr() == avg._flat.untyped_storage().data_ptr()
    assert len({p.grad.data_ptr() for p in m3.parameters()}) == 4
    torch.nn.utils.clip_grad_norm_(m3.parameters(), 0.1)
    for p in m3.parameters():
        p.grad = None
# length-balanced sampler: the two ranks' shares of every global minibatch are disjoint and cover it
import numpy as np
from pykaldi2_b200.data.dataloader import BalancedBatchSampler
L = (np.arange(20) * 37 %% 23 + 5) * 10
bs = BalancedBatchSampler(L, batch_size=3, seed=3)
assert (bs.world, bs.rank) == (world, rank)
mine = [sorted(b) for b in bs]
theirs = [None] * world
torch.distributed.all_gather_object(theirs, mine)
for g in range(len(bs)):
    union = theirs[0][g] + theirs[1][g]
    assert len(union) == 6 and len(set(union)) >= 5          # the padded tail may repeat one utterance
    cost = [56 * max(L[i] for i in t[g]) + sum(L[i] for i in t[g]) for t in theirs]
    assert max(cost) / min(cost) < 1.35
print("rank", rank, "ok")
"""


def test_data_parallel_gloo_world2(tmp_path):
    script = tmp_path / "dp.py"
    script.write_text(_DP_SCRIPT % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_transformer_am_matches_reference_golden():
    """models.transformer.TransformerAM against outputs of the reference's own model on the same
    weights (tests/golden/transformer_golden.npz, written by oracle/make_golden.py from /root/reference):
    state-dict keys are interchangeable; no mask / key-padding mask / key-padding + look-ahead mask."""
    from pykaldi2_b200.models import transformer
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "transformer_golden.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    m = transformer.TransformerAM(8, 16, 2, 32, 2, 0.0, 10)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)
    m.eval()
    x = torch.from_numpy(g["x"])
    kpm = torch.from_numpy(g["kpm"])
    lens = g["lens"]
    with torch.no_grad():
        np.testing.assert_allclose(m(x).numpy(), g["y_nomask"], rtol=1e-4, atol=1e-5)
        y = m(x, None, kpm).numpy()
        y2 = m(x, torch.from_numpy(g["src_mask"]), kpm).numpy()
        y3 = m(x, transformer.look_ahead_mask(x.size(0), 2), kpm).numpy()
    for b, n in enumerate(lens):                     # valid frames (padded frames are never read by the losses)
        np.testing.assert_allclose(y[:n, b], g["y_kpm"][:n, b], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(y2[:n, b], g["y_both"][:n, b], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(y3[:n, b], g["y_both"][:n, b], rtol=1e-4, atol=1e-5)
    # fresh instance: every layer starts from the same weights, as nn.TransformerEncoder's deep copies do
    m2 = transformer.TransformerAM(8, 16, 2, 32, 3, 0.1, 10)
    s2 = m2.state_dict()
    assert torch.equal(s2["transformer.layers.0.encoder_layer.linear1.weight"], s2["transformer.layers.2.encoder_layer.linear1.weight"])


def _make_corpus(tmp_path, n=4, seed=0, label_dim=104):
    """A tiny corpus in the reference's formats: zip of 16-bit wavs + pdf-id / transition-id text files + data yaml."""
    import zipfile
    from pykaldi2_b200.reader import zip_io
    from pykaldi2_b200.data import fbank as fb
    rng = np.random.default_rng(seed)
    wavs, zpath = {}, os.path.join(tmp_path, "train.zip")
    with zipfile.ZipFile(zpath, "w") as z:
        for i in range(n):
            utt = "100-%04d-%04d" % (seed, i)
            x = (0.05 * rng.standard_normal(16000 + 3000 * i)).astype(np.float32)
            wavs[utt] = x
            z.writestr("wav/%s.wav" % utt, zip_io.write_wav(io_bytes(), x))
        z.writestr("README.txt", "not audio")
    pdf, tid = os.path.join(tmp_path, "pdf-ids.txt"), os.path.join(tmp_path, "trans-ids.txt")
    labels = {}
    with open(pdf, "w") as fp, open(tid, "w") as ft:
        for utt, x in wavs.items():
            if utt.endswith("0003"):
                continue                                   # an utterance without labels is dropped
            T = fb.num_frames(len(x)) - 1                  # Kaldi-style label count: one short of our frame count
            lab = rng.integers(0, label_dim, T)
            labels[utt] = lab
            fp.write(utt + " " + " ".join(map(str, lab)) + "\n")
            ft.write(utt + " " + " ".join(map(str, 2 * lab + 1)) + "\n")
    data_yaml = os.path.join(tmp_path, "data.yaml")
    with open(data_yaml, "w") as f:
        f.write("clean_source:\n  1:\n    type: Librispeech\n    wav: %s\n    label: %s\n    aux_label: %s\n" % (zpath, pdf, tid))
    return data_yaml, wavs, labels


def io_bytes():
    import io
    return io.BytesIO()


def _make_kaldi_chain_assets(tmp_path, data_yaml, wavs, P=10, seed=1):
    """Text-form Kaldi inputs of train_chain.py for the corpus of _make_corpus: an alignment model with 3-state
    phones 1..P (phone p, state j -> triple 3(p-1)+j, transition ids 2r+1 = self-loop, 2r+2 = forward), a chain
    model with the 1-state chain topology (pdfs 2(p-1), 2(p-1)+1), a monophone tree, den.fst; the corpus' label
    file is REPLACED by transition-id alignments of the alignment model (plain order).  Returns (ali_dir, chain_dir)."""
    from pykaldi2_b200 import synth
    from pykaldi2_b200.data import fbank as fb
    from pykaldi2_b200.reader import fst_io
    rng = np.random.default_rng(seed)
    lab_path = os.path.join(tmp_path, "ali-tids.txt")
    with open(lab_path, "w") as f:
        for utt, x in wavs.items():
            T = fb.num_frames(len(x)) - 1
            tids, left = [], T
            while left > 0:
                p = int(rng.integers(1, P + 1))
                durs = [int(d) for d in rng.integers(1, 6, size=3)]
                if sum(durs) > left or left - sum(durs) < 3:
                    durs = [1, 1, left - 2] if left >= 3 else None
                if durs is None:                             # fewer than 3 frames left: stretch the last state
                    tids += [tids[-1] - 1] * left            # its self-loop id (plain order ends on the forward id)
                    tids[-left - 1], tids[-1] = tids[-1] - 1, tids[-left - 1]
                    break
                for j, d in enumerate(durs):
                    r = 3 * (p - 1) + j
                    tids += [2 * r + 1] * (d - 1) + [2 * r + 2]
                left -= sum(durs)
            assert len(tids) == T
            f.write(utt + " " + " ".join(map(str, tids)) + "\n")
    with open(data_yaml) as f:
        y = f.read()
    with open(data_yaml, "w") as f:
        f.write(y.replace(os.path.join(tmp_path, "pdf-ids.txt"), lab_path))
    ali_dir, chain_dir = os.path.join(tmp_path, "ali"), os.path.join(tmp_path, "chain")
    os.makedirs(ali_dir); os.makedirs(chain_dir)
    phones = " ".join(str(p) for p in range(1, P + 1))
    with open(os.path.join(ali_dir, "final.mdl.txt"), "w") as f:
        f.write("<TransitionModel>\n<Topology>\n<TopologyEntry>\n<ForPhones> %s </ForPhones>\n" % phones +
                "<State> 0 <PdfClass> 0 <Transition> 0 0.75 <Transition> 1 0.25 </State>\n"
                "<State> 1 <PdfClass> 1 <Transition> 1 0.75 <Transition> 2 0.25 </State>\n"
                "<State> 2 <PdfClass> 2 <Transition> 2 0.75 <Transition> 3 0.25 </State>\n<State> 3 </State>\n"
                "</TopologyEntry>\n</Topology>\n<Triples> %d\n" % (3 * P) +
                "".join("%d %d %d\n" % (p, j, 3 * (p - 1) + j) for p in range(1, P + 1) for j in range(3)) +
                "</Triples>\n<LogProbs> [ 0 ] </LogProbs>\n</TransitionModel>\n")
    with open(os.path.join(chain_dir, "0.trans_mdl.txt"), "w") as f:
        f.write("<TransitionModel>\n<Topology>\n<TopologyEntry>\n<ForPhones> %s </ForPhones>\n" % phones +
                "<State> 0 <ForwardPdfClass> 0 <SelfLoopPdfClass> 1 <Transition> 0 0.5 <Transition> 1 0.5 </State>\n"
                "<State> 1 </State>\n</TopologyEntry>\n</Topology>\n<Tuples> %d\n" % P +
                "".join("%d 0 %d %d\n" % (p, 2 * (p - 1), 2 * (p - 1) + 1) for p in range(1, P + 1)) +
                "</Tuples>\n<LogProbs> [ 0 ] </LogProbs>\n</TransitionModel>\n")
    with open(os.path.join(chain_dir, "tree.txt"), "w") as f:
        f.write("ContextDependency 1 0 ToPdf TE 0 %d ( NULL " % (P + 1) +
                " ".join("TE -1 2 ( CE %d CE %d )" % (2 * (p - 1), 2 * (p - 1) + 1) for p in range(1, P + 1)) +
                " ) EndContextDependency\n")
    fst_io.write_fst_binary(synth.make_den_fst(256, 104, 7, seed=1234), os.path.join(chain_dir, "den.fst"))
    return ali_dir, chain_dir


def test_zip_wav_label_ingestion(tmp_path):
    """the reference's corpus formats (zip of wavs, `utt-id int int ...` label files, data yaml)."""
    import struct
    import yaml
    from pykaldi2_b200.data.speech_dataset import SpeechDataset
    from pykaldi2_b200.data.dataloader import WaveDataloader
    from pykaldi2_b200.reader import zip_io
    data_yaml, wavs, labels = _make_corpus(str(tmp_path))
    with open(data_yaml) as f:
        data = yaml.safe_load(f)
    config = {"data_config": {"load_label": True}, "source_paths": [j for _, j in data["clean_source"].items()]}
    ds = SpeechDataset(config)
    assert len(ds) == 3                                     # 4 wavs, one without labels
    for i in range(len(ds)):
        wav, (utt,), pdf, (tid,) = ds[i]
        ref = np.clip(np.round(wavs[utt].astype(np.float64) * 32768.0), -32768, 32767) / 32768.0   # 16-bit PCM -> [-1, 1)
        assert wav.dtype == np.float32 and np.array_equal(wav, ref.astype(np.float32))
        assert pdf.shape == (len(labels[utt]), 1) and (pdf[:, 0] == labels[utt]).all()
        assert tid.shape == (1, len(labels[utt])) and (tid[0] == 2 * labels[utt] + 1).all()
    batch = next(iter(WaveDataloader(ds, 2)))
    assert set(batch) == {"utt_ids", "wav", "label", "aux"} and len(batch["wav"]) == 2
    # other sample formats: 8 / 24 / 32-bit PCM, float32, stereo, odd-sized chunks before `data`
    x = np.array([0.5, -0.25, 0.125, -1.0])

    def riff(tag, ch, bits, payload, extra=b""):
        fmt = struct.pack("<HHIIHH", tag, ch, 16000, 16000 * ch * bits // 8, ch * bits // 8, bits)
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + extra + b"data" + struct.pack("<I", len(payload)) + payload
        return b"RIFF" + struct.pack("<I", len(body)) + body
    fs, y = zip_io.parse_wav(riff(1, 1, 32, (x * 2 ** 31).clip(-2 ** 31, 2 ** 31 - 1).astype("<i4").tobytes()))
    assert fs == 16000 and np.allclose(y, x.clip(-1, 1 - 2 ** -31), atol=1e-6)
    _, y = zip_io.parse_wav(riff(3, 1, 32, x.astype("<f4").tobytes(), extra=b"LIST" + struct.pack("<I", 3) + b"abc\0"))
    assert np.array_equal(y, x.astype(np.float32))
    v = (x * 2 ** 23).clip(-2 ** 23, 2 ** 23 - 1).astype(np.int32)
    p24 = b"".join(struct.pack("<i", int(t))[:3] for t in v)
    _, y = zip_io.parse_wav(riff(1, 1, 24, p24))
    assert np.allclose(y, x.clip(-1, 1 - 2 ** -23), atol=1e-6)
    _, y = zip_io.parse_wav(riff(1, 2, 16, (x * 32768).clip(-32768, 32767).astype("<i2").tobytes()))
    assert y.shape == (2, 2)
    with pytest.raises(ValueError):
        zip_io.parse_wav(b"RIFFxxxxWAVEjunk")
    with pytest.raises(ValueError):
        zip_io.ZipWaveIO().read_wav(str(tmp_path) + "/a.flac")
    assert zip_io.utt_id_of("/x/y.zip@/wav/100-121669-0001.wav") == "100-121669-0001"


def test_fst_readers_round_trip(tmp_path):
    """SURVEY 8f-2 (reader half): OpenFst binary (vector / standard) and fstprint text forms of a denominator FST."""
    from pykaldi2_b200 import synth
    from pykaldi2_b200.reader import fst_io
    fst = synth.make_den_fst(64, 20, 5, seed=3)
    fst["final"][5] = 0.25
    fst["final"][9] = 0.0
    b = fst_io.write_fst_binary(fst, os.path.join(tmp_path, "den.fst"))
    assert fst_io.FST_MAGIC == struct_unpack_i(b[:4])
    r = fst_io.read_fst(os.path.join(tmp_path, "den.fst"))
    for k in ("src", "dst", "ilabel"):
        assert (r[k] == fst[k]).all(), k
    assert r["num_states"] == 64 and r["start"] == 0
    assert np.array_equal(r["weight"], fst["weight"]) and np.array_equal(r["final"], fst["final"])
    fst_io.write_fst_text(fst, os.path.join(tmp_path, "den.txt"))
    t = fst_io.read_fst(os.path.join(tmp_path, "den.txt"))
    assert (t["src"] == fst["src"]).all() and (t["dst"] == fst["dst"]).all() and (t["ilabel"] == fst["ilabel"]).all()
    np.testing.assert_allclose(t["weight"], fst["weight"], rtol=1e-6)
    assert np.isinf(t["final"][0]) and t["final"][5] == np.float32(0.25) and t["final"][9] == 0.0
    # hand-written fstprint text: start = source of the first line, default weights 0
    h = fst_io.read_fst_text(["2 0 3 3 0.5", "0 1 1 1", "1 2 2 2 1.5", "1"])
    assert h["start"] == 2 and h["num_states"] == 3 and list(h["src"]) == [0, 1, 2] and h["final"][1] == 0.0
    assert h["weight"][list(h["src"]).index(0)] == 0.0
    with pytest.raises(ValueError):
        fst_io.read_fst_binary(b"\\0" * 64)


def struct_unpack_i(b):
    import struct
    return struct.unpack("<i", b)[0]


def test_den_schedule_host_planner():
    """pk2_denfb's host-side schedule (clusters of 8 with work lists + single-CTA kernels on spare SMs): every
    sequence is scheduled exactly once, the single-CTA set is the shortest sequences, loads are balanced."""
    import ctypes as C
    from pykaldi2_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)
    T = np.clip(rng.gamma(6.0, 2.05, 64) * 100 / 3, 50, 1000).astype(np.int32)
    for ncl, spare in ((15, 20), (15, 0), (4, 3), (80, 20)):
        assign = np.full(len(T), -7, np.int32)
        worst = L.pk2_den_plan(T.ctypes.data_as(_lib.vp), len(T), ncl, spare, assign.ctypes.data_as(_lib.vp))
        assert worst > 0 and (assign >= -1).all() and (assign < ncl).all()
        single = assign == -1
        assert single.sum() <= max(spare, 0)
        if single.any():
            assert T[single].max() <= T[~single].min()                  # the shortest ones run on single SMs
            assert T[single].max() * 5.6 <= worst * 1.25                 # and do not outlast the cluster pool by much
        loads = np.array([(T[assign == k] + 8).sum() for k in range(min(ncl, (~single).sum()))])
        assert loads.max() == worst
        if (~single).sum() >= 3 * ncl:
            assert loads.max() <= 1.06 * loads.mean()                    # LPT + local search: within a few percent
    assert L.pk2_den_plan(None, 4, 2, 0, None) == -1


KALDI_TM_TEXT = """<TransitionModel>
<Topology>
<TopologyEntry>
<ForPhones>
2 3
</ForPhones>
<State> 0 <PdfClass> 0 <Transition> 0 0.75 <Transition> 1 0.25 </State>
<State> 1 <PdfClass> 1 <Transition> 1 0.75 <Transition> 2 0.25 </State>
<State> 2 </State>
</TopologyEntry>
<TopologyEntry>
<ForPhones>
1
</ForPhones>
<State> 0 <PdfClass> 0 <Transition> 0 0.5 <Transition> 1 0.25 <Transition> 2 0.25 </State>
<State> 1 <PdfClass> 1 <Transition> 1 0.5 <Transition> 2 0.5 </State>
<State> 2 </State>
</TopologyEntry>
</Topology>
<Triples> 6
1 0 0
1 1 1
2 0 2
2 1 3
3 0 2
3 1 4
</Triples>
<LogProbs>
 [ 0 -0.69 -1.38 -1.38 -0.69 -0.69 -0.28 -1.38 -0.28 -1.38 -0.28 -1.38 -0.28 -1.38 ]
</LogProbs>
</TransitionModel>
"""


def test_kaldi_text_transition_model_and_occs(tmp_path):
    """Transition ids as Kaldi's ComputeDerived enumerates them: tuples in order, one id per topology transition."""
    from pykaldi2_b200 import graphs
    from pykaldi2_b200.reader import kaldi_io
    tm = kaldi_io.read_transition_model_text(KALDI_TM_TEXT)
    #            tid:  1  2  3 | 4  5 | 6  7 | 8  9 | 10 11 | 12 13
    assert list(tm["tid2pdf"]) == [-1, 0, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 4, 4]
    assert list(tm["tid2phone"]) == [0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3]
    assert tm["num_pdfs"] == 5 and tm["phones"] == [1, 2, 3]
    path = os.path.join(tmp_path, "final.mdl.txt")
    with open(path, "w") as f:
        f.write(KALDI_TM_TEXT)
    t = graphs.TidPdfMap.from_kaldi_text(path)
    assert t.num_transition_ids() == 13 and t.transition_id_to_pdf(9) == 3 and t.transition_id_to_phone(12) == 3
    # newer models: <Tuples> with a separate self-loop pdf
    tup = KALDI_TM_TEXT.replace("<Triples> 6", "<Tuples> 6").replace("</Triples>", "</Tuples>")
    tup = re.sub(r"^(\d) (\d) (\d)$", lambda m: "%s %s %s %d" % (m.group(1), m.group(2), m.group(3), int(m.group(3)) + 10), tup, flags=re.M)
    tm2 = kaldi_io.read_transition_model_text(tup)
    assert list(tm2["tid2pdf"][1:6]) == [10, 0, 0, 11, 1]      # self-loop transitions take the self-loop pdf
    with open(os.path.join(tmp_path, "bin.mdl"), "wb") as f:
        f.write(b"\0B<TransitionModel>")
    with pytest.raises(ValueError, match="binary"):
        kaldi_io.read_transition_model_text(os.path.join(tmp_path, "bin.mdl"))
    lp = kaldi_io.log_prior_from_occs(" [ 1 3 4 ]\n")
    np.testing.assert_allclose(np.exp(lp), [0.125, 0.375, 0.5], rtol=1e-6)


def test_alignment_to_supervision_fst():
    """Numerator graph from a frame-level pdf alignment (train_chain.py on a real corpus): the subsampled alignment
    is a path of the FST, every path has T' labels, boundaries move by at most the tolerance."""
    from oracle import chain_ref
    from pykaldi2_b200 import graphs, synth
    rng = np.random.default_rng(2)
    ali = np.repeat(rng.integers(0, 40, 12), rng.integers(3, 15, 12))
    for factor, shift, slack in ((3, 0, 2), (3, 1, 1), (1, 0, 0)):
        sub = ali[shift::factor]
        fst = synth.alignment_to_supervision_fst(ali, factor, shift, slack)
        T = len(sub)
        st = fst["state_times"]
        assert (np.diff(st) >= 0).all() and st[-1] == T and np.isfinite(fst["final"]).sum() == 1
        assert (st[fst["dst"]] == st[fst["src"]] + 1).all()                      # every arc advances one frame
        # the alignment itself is accepted: follow it greedily through the FST
        out = [[] for _ in range(fst["num_states"])]
        for s, d, l in zip(fst["src"], fst["dst"], fst["ilabel"]):
            out[s].append((int(d), int(l) - 1))
        cur = {int(fst["start"])}
        for t in range(T):
            cur = {d for s in cur for d, p in out[s] if p == sub[t]}
            assert cur, "alignment path lost at frame %d" % t
        assert any(np.isfinite(fst["final"][s]) for s in cur)
        if slack == 0:                                                           # no tolerance: exactly one path
            assert len(fst["src"]) == T
        sup = graphs.Supervision(fst, T, 40)
        assert sup.frames_per_sequence == T
        ll = rng.normal(0, 1, (T, 40))
        logz = chain_ref.num_fb_log(ll, fst)[0]
        assert np.isfinite(logz)
    # labels shorter than the features: padded with the last label
    f2 = synth.alignment_to_supervision_fst(ali[:30], 3, 0, 2, n_out=12)
    assert f2["state_times"][-1] == 12


def test_ingestion_directory_source_and_resampling(tmp_path):
    """Directory corpora (no zip), several sources, and the resampling branch of the wav reader."""
    from pykaldi2_b200.data.speech_dataset import SpeechDataset
    from pykaldi2_b200.reader import zip_io
    rng = np.random.default_rng(4)
    srcs = []
    for k in range(2):
        d = os.path.join(tmp_path, "corpus%d" % k, "spk")
        os.makedirs(d)
        lab = os.path.join(tmp_path, "lab%d.txt" % k)
        with open(lab, "w") as f:
            for i in range(2):
                utt = "u%d-%d" % (k, i)
                zip_io.write_wav(os.path.join(d, utt + ".wav"), 0.1 * rng.standard_normal(8000))
                f.write(utt + " " + " ".join(str(j % 7) for j in range(40)) + "\n")
        srcs.append({"type": "Any", "wav": os.path.join(tmp_path, "corpus%d" % k), "label": lab})
    ds = SpeechDataset({"source_paths": srcs, "data_config": {}})
    assert len(ds) == 4 and sorted(ds[i][1][0] for i in range(4)) == ["u0-0", "u0-1", "u1-0", "u1-1"]
    wav, _, pdf, tid = ds[0]
    assert wav.shape == (8000,) and pdf.shape == (40, 1) and tid is None          # no aux_label given
    assert len(SpeechDataset({"source_paths": srcs, "sweep_size": 0.005})) == 1   # sweep_size (hours) caps the epoch
    # 8 kHz file read by a 16 kHz reader: polyphase resampling doubles the length
    p8 = os.path.join(tmp_path, "eight.wav")
    t = np.arange(4000) / 8000.0
    zip_io.write_wav(p8, 0.5 * np.sin(2 * np.pi * 440 * t), fs=8000)
    fs, x = zip_io.ZipWaveIO("float32", 16000).read_wav(p8)
    assert fs == 16000 and abs(len(x) - 8000) <= 1
    ref = 0.5 * np.sin(2 * np.pi * 440 * np.arange(len(x)) / 16000.0)
    assert np.abs(x[200:-200] - ref[200:-200]).max() < 2e-2
    with pytest.raises(ValueError):
        SpeechDataset({"source_paths": []})


# ------------------------------------------------------------------ round 2: sharding, sampling, shuffle buffer ----
def test_balanced_shards_cost_model():
    """dist.balanced_shards: every utterance exactly once, per-rank capacity respected, deterministic, and the modelled
    step time (50 x longest + sum) better balanced than the sorted round-robin deal it replaces."""
    from pykaldi2_b200 import dist as pkdist
    from pykaldi2_b200 import synth
    rng = np.random.default_rng(7)
    for world, per in ((2, 64), (8, 64), (4, 3)):
        L = (synth.make_durations(world * per, rng) * 100).astype(int)
        sh = pkdist.balanced_shards(L, world, per)
        assert sorted(i for s in sh for i in s) == list(range(world * per))
        assert all(len(s) == per for s in sh)
        assert sh == pkdist.balanced_shards(L, world, per)
        cost = [50 * L[s].max() + L[s].sum() for s in sh]
        srt = np.sort(L)[::-1]
        rr = [50 * srt[r::world].max() + srt[r::world].sum() for r in range(world)]
        assert max(cost) <= max(rr)
        if per >= 64:
            assert max(cost) / np.mean(cost) < 1.03
    # ragged tail: fewer utterances than world * per_rank
    sh = pkdist.balanced_shards([5, 9, 2], 2)
    assert sorted(i for s in sh for i in s) == [0, 1, 2] and max(len(s) for s in sh) == 2
    with pytest.raises(ValueError):
        pkdist.balanced_shards([1, 2, 3], 2, 1)


def test_balanced_batch_sampler_covers_epoch_and_reseeds():
    from pykaldi2_b200.data.dataloader import BalancedBatchSampler
    L = np.arange(1, 41) * 10
    seen = []
    per_rank = []
    for r in range(4):
        bs = BalancedBatchSampler(L, batch_size=3, world=4, rank=r, seed=5)
        assert len(bs) == 4                                   # ceil(40 / 12)
        batches = list(bs)
        assert all(len(b) == 3 for b in batches)
        per_rank.append(batches)
        seen += [i for b in batches for i in b]
    assert set(seen) == set(range(40)) and len(seen) == 48    # tail padded by wrap-around
    # the four ranks' shares of one global batch are disjoint
    g0 = [i for r in range(4) for i in per_rank[r][0]]
    assert len(set(g0)) == 12
    bs = BalancedBatchSampler(L, batch_size=3, world=4, rank=0, seed=5)
    e0 = list(bs)
    bs.set_epoch(1)
    assert list(bs) != e0
    bs.set_epoch(0)
    assert list(bs) == e0


def test_chunk_pool_draws_every_chunk_once():
    """pipeline.ChunkPool (the reference's DataBuffer, data/sr_dataset.py:55-84): random draws, nothing lost."""
    from pykaldi2_b200 import pipeline
    pool = pipeline.ChunkPool(50, 4, 3, torch.device("cpu"), seed=1)
    ids = torch.arange(37, dtype=torch.float32)
    pool.add(ids[:20].view(-1, 1, 1).expand(-1, 4, 3), ids[:20].long().view(-1, 1).expand(-1, 4))
    out = []
    x, y = pool.draw(8)
    assert x.shape == (8, 4, 3) and y.shape == (8, 4) and pool.n == 12
    out += x[:, 0, 0].tolist()
    assert (x[:, 0, 0].long() == y[:, 0]).all()
    pool.add(ids[20:].view(-1, 1, 1).expand(-1, 4, 3), ids[20:].long().view(-1, 1).expand(-1, 4))
    while pool.n:
        x, y = pool.draw(8)
        assert (x[:, 0, 0].long() == y[:, 0]).all()
        out += x[:, 0, 0].tolist()
    assert sorted(out) == list(range(37))
    assert out != sorted(out)                                  # shuffled
    with pytest.raises(RuntimeError):
        pool.add(torch.zeros(51, 4, 3), torch.zeros(51, 4, dtype=torch.long))


def test_speech_dataset_epoch_semantics(tmp_path):
    """Sequence mode visits every utterance whatever sweep_size says (reference data/sr_dataset.py:192-194); chunk mode
    draws a fresh random subset per epoch instead of a fixed prefix."""
    from pykaldi2_b200.data.speech_dataset import SpeechDataset
    from pykaldi2_b200.reader import zip_io
    rng = np.random.default_rng(0)
    d = os.path.join(tmp_path, "c")
    os.makedirs(d)
    lab = os.path.join(tmp_path, "lab.txt")
    with open(lab, "w") as f:
        for i in range(12):
            utt = "u%02d" % i
            zip_io.write_wav(os.path.join(d, utt + ".wav"), 0.1 * rng.standard_normal(1600 + 160 * i))
            f.write(utt + " " + " ".join("1" for _ in range(8 + i)) + "\n")
    srcs = [{"type": "Any", "wav": d, "label": lab}]
    seq = SpeechDataset({"source_paths": srcs, "sweep_size": 0.01, "data_config": {"sequence_mode": True}})
    assert len(seq) == 12
    assert seq.utt_lengths().tolist() == [8 + i for i in range(12)]
    ch = SpeechDataset({"source_paths": srcs, "sweep_size": 0.01, "data_config": {"sequence_mode": False}})
    assert len(ch) == 2                                        # 0.01 h at 12.3 s per utterance
    picks = set()
    for ep in range(8):
        ch.set_epoch(ep)
        picks |= {ch[i][1][0] for i in range(len(ch))}
    assert len(picks) > 4                                      # not the same prefix every epoch
    assert ch.reader._zips == {}                               # no handle survives the constructor (fork safety)


def test_lattice_frame_acc_matches_oracle_per_arc():
    """graphs.Lattice.frame_acc (per-transition-id tables, one pass over the out-arcs, stored in<-out permutation)
    against the oracle's per-arc definition, both criteria, both silence conventions; in-arc and out-arc order."""
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    rng = np.random.default_rng(19)
    N, T = 40, 17
    lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=5, kmax=11, eps_frac=0.1)
    tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 3 + 1, 0)
    L = graphs.Lattice(lat)
    t_out = np.repeat(L.state_time[:L.num_states], np.diff(L.out_off))
    t_in = np.repeat(L.state_time[:L.num_states], np.diff(L.in_off)) - 1
    for criterion in ("smbr", "mpfe"):
        for sil in ([1, 2], [], [3]):
            a_in, a_out = L.frame_acc(ali, tid2pdf, tid2phone, criterion, sil)
            ref_out = [lattice_ref.mpe_frame_acc(int(t), int(ali[tt]), tid2pdf, tid2phone, criterion, set(sil))
                       for t, tt in zip(L.out_tid, t_out)]
            ref_in = [lattice_ref.mpe_frame_acc(int(t), int(ali[tt]), tid2pdf, tid2phone, criterion, set(sil))
                      for t, tt in zip(L.in_tid, t_in)]
            assert a_out.dtype == np.uint8 and (a_out == np.asarray(ref_out, np.uint8)).all()
            assert (a_in == np.asarray(ref_in, np.uint8)).all()
    with pytest.raises(ValueError):
        L.frame_acc(ali[:-1], tid2pdf, tid2phone, "smbr", [1])


def test_train_chain_data_side_with_kaldi_assets(tmp_path):
    """The host half of train_chain.py -ali_dir -chain_dir without a GPU: asset loading, the loader's batch_transform
    (numerator graphs built where the batch is collated, also in worker processes), graphs.Supervision objects."""
    import importlib.util
    from types import SimpleNamespace
    import yaml
    from pykaldi2_b200 import graphs
    from pykaldi2_b200.data import fbank as fb
    from pykaldi2_b200.data.dataloader import WaveDataloader
    from pykaldi2_b200.data.speech_dataset import SpeechDataset
    sys.path.insert(0, os.path.join(ROOT, "bin"))
    spec = importlib.util.spec_from_file_location("train_chain_mod", os.path.join(ROOT, "bin", "train_chain.py"))
    tc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tc)
    data_yaml, wavs, _ = _make_corpus(str(tmp_path), n=5, seed=5)
    ali_dir, chain_dir = _make_kaldi_chain_assets(str(tmp_path), data_yaml, wavs)
    kaldi = tc.load_kaldi_assets(SimpleNamespace(ali_dir=ali_dir, chain_dir=chain_dir))
    assert kaldi["tree"].context_width() == 1 and kaldi["chain_tm"]["num_pdfs"] == 20
    opts = tc.SupervisionOptions()
    assert (opts.left_tolerance, opts.right_tolerance, opts.frame_subsampling_factor) == (5, 5, 3)

    def transform(batch):
        batch["sup"] = [tc.kaldi_supervision(kaldi, opts, lab[:min(fb.num_frames(len(w)), len(lab)), 0])
                        for w, lab in zip(batch["wav"], batch["label"])]
        return batch
    with open(data_yaml) as f:
        data = yaml.safe_load(f)
    cfg = {"source_paths": [v for v in data["clean_source"].values()], "data_config": {"sequence_mode": True}}
    ds = SpeechDataset(cfg)
    for workers in (0, 2):
        seen = 0
        for batch in WaveDataloader(ds, 2, num_workers=workers, batch_transform=transform):
            for w, lab, (fst, t_sub) in zip(batch["wav"], batch["label"], batch["sup"]):
                n = min(fb.num_frames(len(w)), len(lab))
                assert t_sub == (n - 1) // 3 + 1
                sup = graphs.Supervision(fst, t_sub, 104)
                assert sup.frames_per_sequence == t_sub
                assert int(fst["ilabel"].max()) <= 20 and int(fst["ilabel"].min()) >= 1
                seen += 1
        assert seen == 4                                    # the utterance without aux labels is dropped by the dataset


def test_lattice_and_supervision_batches_on_cpu(monkeypatch):
    """The host half of LatticeBatch / SupervisionBatch without a GPU (the upload is replaced by plain tensors): state
    times and drop masks against the oracle with and without epsilon arcs, index validation, int32 batch arrays."""
    from oracle import lattice_ref
    from pykaldi2_b200 import graphs, synth
    monkeypatch.setattr(graphs, "_upload",
                        lambda host, device: ({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in host.items()}, None))
    rng = np.random.default_rng(11)
    for eps in (0.0, 0.1):
        N, Ts = 60, [17, 9, 12]
        lats, alis, olat = [], [], []
        for T in Ts:
            lat, tid2pdf, ali = synth.make_lattice(T, N, rng, kmin=4, kmax=9, ali_drop=0.2, eps_frac=eps)
            olat.append(lat); alis.append(ali); lats.append(graphs.Lattice(lat))
        lb = graphs.LatticeBatch(lats, tid2pdf, alis, device="cpu")
        pred = rng.normal(0, 3.0, (len(Ts), max(Ts), N)).astype(np.float32)
        for b, T in enumerate(Ts):
            _, _, drop, times = lattice_ref.lattice_fb_mmi(pred[b, :T], olat[b], tid2pdf, alis[b])
            assert (lats[b].state_times_orig == times).all()
            assert (lb.keep_host[b] == (~drop).astype(np.uint8)).all()
        assert lb.total_frames == sum(Ts) and lb.total_arcs == sum(len(l.in_src) for l in lats)
        assert lb.max_pdf < N and lb._dev["out_dst"].dtype == torch.int32 and lb._dev["in_src"].dtype == torch.int32
        # concatenated state indices point into the right lattice
        off = np.concatenate([[0], np.cumsum([l.num_states for l in lats])])
        od = lb._dev["out_dst"].numpy()
        k = 0
        for b, l in enumerate(lats):
            seg = od[k:k + len(l.out_dst)]
            assert seg.min() >= off[b] and seg.max() < off[b + 1]
            k += len(l.out_dst)
    with pytest.raises(ValueError):
        graphs.LatticeBatch(lats, tid2pdf[:10], alis, device="cpu")
    sups = [graphs.Supervision(synth.make_supervision_fst(T, 50, rng), T, 50) for T in (5, 9, 1)]
    sb = graphs.SupervisionBatch(sups, device="cpu")
    assert sb.n_seq == 3 and sb.num_frames_host == [5, 9, 1] and sb._dev["out_dst"].dtype == torch.int32
