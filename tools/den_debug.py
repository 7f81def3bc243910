#!/usr/bin/env python
"""Profiling aid: time the denominator kernels with parts switched off (PK2_DEN_DEBUG)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykaldi2_b200 import graphs, synth
from pykaldi2_b200.ops import ops
from tools.kernel_bench import timeit
N, S = 5768, 8192
dev = torch.device("cuda", 0)
den = graphs.DenominatorGraph(synth.make_den_fst(S, N, 7, seed=1234), N)
opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4)
rng = np.random.default_rng(0)
for Bn, Tu in ((16, 200), (64, 200)):
    sup = [graphs.Supervision(synth.make_supervision_fst(Tu, N, rng), Tu, N) for _ in range(Bn)]
    sb = graphs.SupervisionBatch(sup, device=dev)
    pred = torch.randn(Bn, Tu, N, device=dev) * 2
    for K in (1, 2, 4):
        if Bn * K > 148:
            continue
        ms = timeit(lambda: ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=K), iters=3, warm=1)
        print(json.dumps({"debug": os.environ.get("PK2_DEN_DEBUG", "0"), "B": Bn, "K": K, "us_per_frame": 1e3 * ms / Tu}))
