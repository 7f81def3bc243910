"""Integration tests on the GPU box: the three trainers run a few optimizer steps on synthetic data
through the reference-compatible CLI (flags of bin/train_ce.py:46-62, bin/train_se.py:56-75,
bin/train_chain.py:61-83), write reference-named checkpoints that reload, and the chain loss of a
trained step decreases."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, args, tmp_path, timeout=600):
    cmd = [sys.executable, os.path.join(ROOT, "bin", script)] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_train_ce_synthetic(tmp_path):
    out = run("train_ce.py", ["-exp_dir", str(tmp_path), "-train_config", "configs/ce_test.yaml", "-batch_size", "16",
                              "-synthetic", "12", "-print_freq", "1", "-max_steps", "3", "-lr", "0.001",
                              "-global_mvn", "true"], tmp_path)
    assert "Epoch: [0]" in out and "iRTF" in out
    ck = torch.load(os.path.join(tmp_path, "model.0.tar"), map_location="cpu")
    assert set(ck) == {"model", "optimizer", "epoch"} and "lstm.weight_ih_l0" in ck["model"]
    assert os.path.exists(os.path.join(tmp_path, "transform.pkl"))
    # resume from the checkpoint (reference flag -resume_from_model)
    run("train_ce.py", ["-exp_dir", str(tmp_path), "-train_config", "configs/ce_test.yaml", "-batch_size", "16",
                        "-synthetic", "12", "-max_steps", "1", "-num_epochs", "1",
                        "-resume_from_model", os.path.join(tmp_path, "model.0.tar")], tmp_path)


def test_train_chain_synthetic(tmp_path):
    out = run("train_chain.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-batch_size", "4",
                                 "-synthetic", "16", "-den_states", "256", "-print_freq", "1", "-lr", "0.01",
                                 "-warmup_steps", "2", "-max_steps", "4"], tmp_path)
    assert out.count("Epoch: [0]") >= 3
    # a denominator graph from a file (OpenFst binary as Kaldi writes den.fst) instead of the synthetic one
    from pykaldi2_b200 import synth
    from pykaldi2_b200.reader import fst_io
    den_path = os.path.join(tmp_path, "den.fst")
    fst_io.write_fst_binary(synth.make_den_fst(256, 104, 7, seed=1234), den_path)
    o0 = run("train_chain.py", ["-exp_dir", str(tmp_path / "d"), "-config", "configs/ce_test.yaml", "-batch_size", "4",
                                "-synthetic", "4", "-den_fst", den_path, "-print_freq", "1", "-max_steps", "1"], tmp_path)
    assert "Epoch: [0]" in o0
    ck = torch.load(os.path.join(tmp_path, "chain.model.0.tar"), map_location="cpu")
    assert "output_layer.weight" in ck["model"]
    # per-utterance calling pattern of the reference gives the same first-step loss as the batched call
    o1 = run("train_chain.py", ["-exp_dir", str(tmp_path / "a"), "-config", "configs/ce_test.yaml", "-batch_size", "4",
                                "-synthetic", "4", "-den_states", "256", "-print_freq", "1", "-max_steps", "1",
                                "-per_utt_loss", "1"], tmp_path)
    o2 = run("train_chain.py", ["-exp_dir", str(tmp_path / "b"), "-config", "configs/ce_test.yaml", "-batch_size", "4",
                                "-synthetic", "4", "-den_states", "256", "-print_freq", "1", "-max_steps", "1"], tmp_path)
    # meters read one step late (host one step ahead of the GPU): same training, lagging printout
    o3 = run("train_chain.py", ["-exp_dir", str(tmp_path / "c"), "-config", "configs/ce_test.yaml", "-batch_size", "4",
                                "-synthetic", "12", "-den_states", "256", "-print_freq", "1", "-max_steps", "3",
                                "-async_meters", "1"], tmp_path)
    assert o3.count("Epoch: [0]") >= 3
    l1 = [l for l in o1.splitlines() if l.startswith("Epoch")][0].split("Loss")[1].split()[0]
    l2 = [l for l in o2.splitlines() if l.startswith("Epoch")][0].split("Loss")[1].split()[0]
    assert abs(float(l1) - float(l2)) <= 1e-3 * abs(float(l2)) + 1e-6


def test_train_se_synthetic(tmp_path):
    out = run("train_se.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-batch_size", "2",
                              "-synthetic", "6", "-print_freq", "1", "-lr", "0.0001", "-max_steps", "2"], tmp_path)
    assert out.count("Epoch: [0]") >= 2
    assert os.path.exists(os.path.join(tmp_path, "model.se.0.tar"))
    out2 = run("train_se.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-batch_size", "2",
                               "-synthetic", "2", "-print_freq", "1", "-max_steps", "1", "-batched_loss", "0"], tmp_path)
    assert "Epoch: [0]" in out2
    # transition model / priors from Kaldi text files (-trans_model, -prior_path) instead of the synthetic maps:
    # one emitting state with a self-loop and a forward transition per pdf gives the ids 2p+1, 2p+2 the synthetic
    # alignments use
    N = 104
    mdl = ["<TransitionModel>", "<Topology>", "<TopologyEntry>", "<ForPhones>", " ".join(str(i) for i in range(1, N // 3 + 2)),
           "</ForPhones>", "<State> 0 <PdfClass> 0 <Transition> 0 0.5 <Transition> 1 0.5 </State>", "<State> 1 </State>",
           "</TopologyEntry>", "</Topology>", "<Triples> %d" % N] + ["%d 0 %d" % (p // 3 + 1, p) for p in range(N)] + \
          ["</Triples>", "<LogProbs>", " [ 0 " + " ".join(["-0.69"] * (2 * N)) + " ]", "</LogProbs>", "</TransitionModel>"]
    with open(os.path.join(tmp_path, "final.mdl.txt"), "w") as f:
        f.write("\n".join(mdl) + "\n")
    with open(os.path.join(tmp_path, "final.occs.txt"), "w") as f:
        f.write(" [ " + " ".join(str(10 + (i % 7)) for i in range(N)) + " ]\n")
    out4 = run("train_se.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-batch_size", "2",
                               "-synthetic", "2", "-print_freq", "1", "-max_steps", "1", "-criterion", "smbr",
                               "-trans_model", os.path.join(tmp_path, "final.mdl.txt"),
                               "-prior_path", os.path.join(tmp_path, "final.occs.txt")], tmp_path)
    assert "Epoch: [0]" in out4
    # -criterion switch (reference bin/train_se.py:62,216-219), the reference's per-utterance calling pattern
    out3 = run("train_se.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-batch_size", "2",
                               "-synthetic", "2", "-print_freq", "1", "-max_steps", "1", "-criterion", "mpfe",
                               "-batched_loss", "0"], tmp_path)
    assert "Epoch: [0]" in out3


def test_trainers_on_zip_corpus(tmp_path):
    """train_ce.py / train_se.py on a corpus in the reference's formats (zip of wavs + label text
    files + data yaml) instead of -synthetic."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_host import _make_corpus
    data_yaml, wavs, labels = _make_corpus(str(tmp_path), n=6, seed=3)
    out = run("train_ce.py", ["-exp_dir", str(tmp_path), "-train_config", "configs/ce_test.yaml", "-data_config", data_yaml,
                              "-batch_size", "4", "-print_freq", "1", "-lr", "0.001", "-max_steps", "2"], tmp_path)
    assert out.count("Epoch: [0]") >= 1 and os.path.exists(os.path.join(tmp_path, "model.0.tar"))
    out = run("train_se.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-data", data_yaml,
                              "-batch_size", "2", "-print_freq", "1", "-max_steps", "1"], tmp_path)
    assert "Epoch: [0]" in out
    # LF-MMI on the same corpus: denominator graph from a file, numerator graphs from the pdf alignments
    from pykaldi2_b200 import synth
    from pykaldi2_b200.reader import fst_io
    den_path = os.path.join(tmp_path, "den.fst")
    fst_io.write_fst_binary(synth.make_den_fst(256, 104, 7, seed=1234), den_path)
    out = run("train_chain.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-data", data_yaml,
                                 "-den_fst", den_path, "-batch_size", "2", "-print_freq", "1", "-max_steps", "2"], tmp_path)
    assert out.count("Epoch: [0]") >= 2


def test_train_chain_with_kaldi_text_assets(tmp_path):
    """SURVEY 8f-2 end to end: train_chain.py -ali_dir / -chain_dir with a text-form alignment transition model
    (3-state phones), chain transition model (1-state chain topology), monophone tree and den.fst; the label file
    holds the alignment model's transition ids, as in the reference (bin/train_chain.py:162-181, 262-272)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_host import _make_corpus, _make_kaldi_chain_assets
    data_yaml, wavs, _ = _make_corpus(str(tmp_path), n=5, seed=5)
    ali_dir, chain_dir = _make_kaldi_chain_assets(str(tmp_path), data_yaml, wavs)
    out = run("train_chain.py", ["-exp_dir", str(tmp_path), "-config", "configs/ce_test.yaml", "-data", data_yaml,
                                 "-ali_dir", ali_dir, "-chain_dir", chain_dir, "-batch_size", "2", "-print_freq", "1",
                                 "-max_steps", "2"], tmp_path)
    assert out.count("Epoch: [0]") >= 2
    loss = [float(l.split("Loss")[1].split()[0]) for l in out.splitlines() if l.startswith("Epoch")]
    assert all(np.isfinite(v) for v in loss)
