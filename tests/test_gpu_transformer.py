"""Transformer AM path on the GPU: bf16 autocast forward against the fp32 CPU forward of the same
weights (which tests/test_host.py pins against the reference's own model), MMI loss through it, and the trainer."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_transformer_am_bf16_close_to_fp32_and_trains():
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.models import transformer
    from pykaldi2_b200.ops import ops
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    F_, D, N = 80, 128, 96
    m = transformer.TransformerAM(F_, D, 4, 256, 3, 0.0, N)
    m.eval()
    T, lens = 50, [50, 37, 21]
    x = torch.randn(T, len(lens), F_)
    kpm = torch.arange(T)[None, :] >= torch.tensor(lens)[:, None]
    mask = transformer.look_ahead_mask(T, 5)
    with torch.no_grad():
        ref = m(x, mask, kpm)                                   # fp32 on the CPU
        mg = m.to(dev)
        out = mg(x.to(dev), mask.to(dev), kpm.to(dev)).cpu()    # bf16 autocast on the GPU
        mg.autocast_bf16 = False
        out32 = mg(x.to(dev), mask.to(dev), kpm.to(dev)).cpu()
        mg.autocast_bf16 = True
    for b, n in enumerate(lens):
        err32 = (out32[:n, b] - ref[:n, b]).norm() / ref[:n, b].norm()
        assert float(err32) < 3e-3, float(err32)                # fp32 parameters; cuDNN runs the Conv1d in TF32
        err = (out[:n, b] - ref[:n, b]).norm() / ref[:n, b].norm()
        assert float(err) < 3e-2, float(err)                    # bf16 operands, fp32 accumulate
    # one MMI step through the model (batched loss, as bin/train_transformer_se.py does)
    mg.train()
    rng = np.random.default_rng(0)
    lats, alis = [], []
    for n in lens:
        lat, tid2pdf, ali = synth.make_lattice(n, N, rng, kmin=4, kmax=8)
        lats.append(graphs.Lattice(lat)); alis.append(ali)
    lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev)
    pred = mg(x.to(dev), None, kpm.to(dev)).transpose(0, 1).contiguous()
    loss = ops.MMIFunction.apply_batch(pred, lb)
    loss.backward()
    g = mg.transformer.layers[0].encoder_layer.self_attn.in_proj_weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0


def test_train_transformer_se_synthetic(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, os.path.join(ROOT, "bin", "train_transformer_se.py"), "-exp_dir", str(tmp_path),
           "-config", os.path.join(ROOT, "configs", "ce_test.yaml"), "-batch_size", "2", "-synthetic", "4",
           "-print_freq", "1", "-lr", "0.0001", "-max_steps", "2", "-dim_model", "64", "-nheads", "2",
           "-ff_size", "128", "-nlayers", "2", "-look_ahead", "10", "-criterion", "smbr"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("Epoch: [0]") >= 2
    assert os.path.exists(os.path.join(tmp_path, "model.se.0.tar"))
