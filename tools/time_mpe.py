#!/usr/bin/env python
"""Timing of the lattice-MMI and sMBR kernels on a C3-shaped batch (4 utterances, ~250 arcs per frame)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t_start = time.time()
import numpy as np, torch
from pykaldi2_b200 import graphs, synth
from pykaldi2_b200.ops import ops
dev = torch.device('cuda', 0)
print('imports', time.time() - t_start, flush=True)
rng = np.random.default_rng(3)
N = 5768
Ts = [1500, 1200, 900, 600]
lats, alis = [], []
for T in Ts:
    ali = rng.integers(1, 2 * N, T).astype(np.int32)
    lat, tid2pdf, _ = synth.make_lattice(T, N, rng, num_ali=ali)
    lats.append(graphs.Lattice(lat)); alis.append(ali)
tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 3 + 1, 0)
print('lattices', time.time() - t_start, flush=True)
pred = torch.randn(len(Ts), max(Ts), N, device=dev)
lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev, mpe=("smbr", tid2phone, [1]))
for name, fn in (("mmi", lambda: ops.lattice_mmi(pred, lb)), ("mpe", lambda: ops.lattice_mpe(pred, lb))):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    print(name, "ms per call", (time.time() - t0) / 3 * 1e3, flush=True)
