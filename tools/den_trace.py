#!/usr/bin/env python
"""Per-phase clock64 trace of the denominator forward/backward frame loops (frames 64..71 of cluster 0)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykaldi2_b200 import _lib, graphs, synth
from pykaldi2_b200.ops import ops
N, S = 5768, 8192
dev = torch.device("cuda", 0)
L = _lib.lib()
den = graphs.DenominatorGraph(synth.make_den_fst(S, N, 7, seed=1234), N)
opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4)
rng = np.random.default_rng(0)
Bn, Tu = 32, 200
sup = [graphs.Supervision(synth.make_supervision_fst(Tu, N, rng), Tu, N) for _ in range(Bn)]
sb = graphs.SupervisionBatch(sup, device=dev)
pred = torch.randn(Bn, Tu, N, device=dev) * 2
buf = torch.zeros(128, dtype=torch.int64, device=dev)
fn = ["top", "pass done", "exchange done", "leaky done", "cp wait+sync", "exp done", "sync+prefetch"]
bn = ["top", "cp wait+sync", "exp+sync+prefetch", "beta pass", "gamma pass", "exchange done", "grad store"]
Ks = [int(k) for k in sys.argv[1:]] or [1, 2, 4, 8]
fn8 = ["top", "alpha pass", "sync+send", "rows landed", "sum", "-", "-"]
bn8 = ["top", "beta pass", "sync+send", "gamma pass", "sync+stores", "rows landed", "-"]
for K in Ks:
    if K == 8:
        fn, bn = fn8, bn8
    ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=K)
    L.pk2_den_set_profile_buffer(_lib.ptr(buf))
    ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=K)
    torch.cuda.synchronize()
    L.pk2_den_set_profile_buffer(None)
    t = buf.cpu().view(2, 8, 8).numpy()
    print("K=%d forward (cycles after frame top):" % K)
    for f in (5, 6):
        print("  frame %d:" % (64 + f), "  ".join("%s +%d" % (fn[e], t[0][f][e] - t[0][f][0]) for e in range(1, 5 if K == 8 else 7)),
              " period %d" % (t[0][f + 1][0] - t[0][f][0]))
    print("K=%d backward:" % K)
    for f in (5, 6):      # backward runs t downwards: frame f+1 precedes frame f
        print("  frame %d:" % (64 + f), "  ".join("%s +%d" % (bn[e], t[1][f][e] - t[1][f][0]) for e in range(1, 6 if K == 8 else 7)),
              " period %d" % (t[1][f - 1][0] - t[1][f][0]))
