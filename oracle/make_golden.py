"""Generate tests/golden/*.npz by running the REFERENCE's own Python code.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
The reference is imported read-only with two shims:
a stub ``soundfile`` module (reader/reader.py:28-40 would otherwise shell out to
apt/pip) and ``np.int = int`` (simulation/freq_analysis.py:66 uses the removed alias).
The .npz fixtures are committed; nothing on the GPU box reads /root/reference.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)
    import data.sr_dataset as srd          # noqa: E402
    import reader.preprocess as prep       # noqa: E402
    import data.dataloader as dl           # noqa: E402
    return srd, prep, dl


def main():
    srd, prep, dl = import_reference()
    os.makedirs(OUT, exist_ok=True)
    gen = srd.DataGeneratorTrain.__new__(srd.DataGeneratorTrain)
    gen._window_file = "mel80_window.txt"
    gen._gen_window()
    np.save(os.path.join(OUT, "mel80_window.npy"), gen._window.astype(np.float32))

    rng = np.random.default_rng(20260925)
    out = {}
    for i, n in enumerate([560, 401, 16000, 24123, 40000]):
        wav = (0.05 * rng.standard_normal(n)).astype(np.float32)
        if i == 4:  # a louder, tonal signal as well
            t = np.arange(n) / 16000.0
            wav = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.01 * rng.standard_normal(n)).astype(np.float32)
        fb = gen._logfbank_extractor(wav)
        out["wav%d" % i] = wav
        out["fbank%d" % i] = fb.astype(np.float32)
        out["cmn%d" % i] = prep.cmn(fb, axis=0).astype(np.float32)
    # global MVN
    tr = prep.GlobalMeanVarianceNormalization()
    tr.mean_vec = rng.normal(0, 1, (1, 80)).astype(np.float32)
    tr.std_vec = np.abs(rng.normal(1, 0.2, (1, 80))).astype(np.float32)
    out["mvn_mean"], out["mvn_std"] = tr.mean_vec, tr.std_vec
    out["mvn4"] = tr.apply_on_ndarray(out["cmn4"]).astype(np.float32)
    # chunking (data/sr_dataset.py:40-52)
    segs = srd._utt2seg(out["cmn4"].T, 80, 80)
    out["seg4"] = np.stack([s.T for s in segs]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "fbank_golden.npz"), **out)

    # collate functions (data/dataloader.py:55-63, :94-136)
    chunk = dl.ChunkDataloader.__new__(dl.ChunkDataloader)
    seq = dl.SeqDataloader.__new__(dl.SeqDataloader)
    seq.test_only = False
    items = []
    for j, T in enumerate([7, 4, 9]):
        feats = rng.normal(0, 1, (T, 80)).astype(np.float32)
        lab = rng.integers(0, 100, (T, 1))
        aux = [rng.integers(1, 200, (1, T))]
        items.append((feats, ["utt%d" % j], lab, aux))
    b = seq.collate_fn(items)
    col = {"x": b["x"].numpy(), "y": b["y"].numpy(), "num_frs": np.asarray(b["num_frs"])}
    for j, it in enumerate(items):
        col["feat%d" % j], col["lab%d" % j], col["aux%d" % j] = it[0], it[2], it[3][0]
    np.savez_compressed(os.path.join(OUT, "collate_golden.npz"), **col)
    print("golden vectors written to", OUT)



def transformer_golden():
    """Reference models/transformer.py::TransformerAM (stock nn.TransformerEncoder) on CPU fp32: weights,
    input, masks and output of a tiny instance -> tests/golden/transformer_golden.npz."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("ref_transformer", os.path.join(REF, "models", "transformer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(7)
    m = mod.TransformerAM(8, 16, 2, 32, 2, 0.0, 10)
    with torch.no_grad():                      # the reference's layers are deep copies: make them differ
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
    m.eval()
    T, B = 9, 3
    x = torch.randn(T, B, 8)
    lens = [9, 6, 4]
    kpm = torch.ones(B, T)
    for i, n in enumerate(lens):
        kpm[i, :n] = 0
    kpm = kpm.bool()
    keep = torch.tril(torch.ones(T, T), diagonal=2)
    src_mask = keep.float().masked_fill(keep == 0, float("-inf")).masked_fill(keep == 1, 0.0)
    out = {"x": x.numpy(), "kpm": kpm.numpy(), "src_mask": src_mask.numpy(), "lens": np.array(lens)}

    def run(x, src_mask=None, kpm=None):
        # TransformerAM.forward (models/transformer.py:86-93) with nn.TransformerEncoder.forward written out as
        # what it was in the torch the reference targets (1.2): a loop over the layers and the final norm.  The
        # container of torch 2.11 probes `layers[0].self_attn` for its fused fast path, which the reference's
        # wrapper layer (TransformerEncoderLayerWithConv1d) does not have; the layers themselves are the reference's.
        h = m.input_layer(x)
        for layer in m.transformer.layers:
            h = layer(h, src_mask, kpm)
        return m.output_layer(m.transformer.norm(h))

    with torch.no_grad():
        out["y_nomask"] = run(x).numpy()
        out["y_kpm"] = run(x, None, kpm).numpy()
        out["y_both"] = run(x, src_mask, kpm).numpy()
    for k, v in m.state_dict().items():
        out["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "transformer_golden.npz"), **out)
    print("transformer_golden.npz:", {k: v.shape for k, v in out.items() if not k.startswith("sd/")})


if __name__ == "__main__":
    main()
    transformer_golden()
