#!/usr/bin/env python
"""Randomised consistency check of the denominator schedules: automatic (clusters of 8 with work lists, two-slot
forward, single-CTA kernels on spare SMs) against the streaming kernels, over random batch sizes and lengths."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykaldi2_b200 import graphs, synth
from pykaldi2_b200.ops import ops

dev = torch.device("cuda", 0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
dens = {}
worst = 0.0
t0 = time.time()
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    S, N = [(256, 52), (2048, 400), (1024, 128)][trial % 3]
    if (S, N) not in dens:
        dens[(S, N)] = graphs.DenominatorGraph(synth.make_den_fst(S, N, 7, seed=S), N)
    den = dens[(S, N)]
    B = int(rng.integers(1, 72))
    Ts = [int(t) for t in rng.integers(1, 90, size=B)]
    if trial % 4 == 0:
        Ts[int(rng.integers(0, B))] = 1
    sups = [graphs.Supervision(synth.make_supervision_fst(T, N, rng), T, N) for T in Ts]
    sb = graphs.SupervisionBatch(sups, device=dev)
    pred = torch.from_numpy(rng.normal(0, 2.0, (B, max(Ts), N)).astype(np.float32)).to(dev)
    o0, g0 = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=0)
    o8, g8 = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=8)
    o1, g1 = ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=int(rng.choice([1, 2, 4])))
    torch.cuda.synchronize()
    for o, g in ((o0, g0), (o8, g8)):
        assert torch.allclose(o, o1, rtol=1e-5), (trial, B)
        err = float((g - g1).abs().max())
        worst = max(worst, err)
        assert torch.allclose(g, g1, rtol=1e-3, atol=1e-6), (trial, B, err)
    print("trial %d: S=%d B=%d max|dgrad|=%.2e ok" % (trial, S, B, float((g0 - g1).abs().max())), flush=True)
print("all ok, worst abs diff %.3e, %.1f s" % (worst, time.time() - t0))
