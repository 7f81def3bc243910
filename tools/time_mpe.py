#!/usr/bin/env python
"""Timing of the lattice-MMI and sMBR kernels on a C3-shaped batch (BASELINE config 3: 4 utterances per GPU,
K_t ~ U{32..96} states per frame, ~250 arcs per frame, N = 5768).  CUDA events, warm.  (The round-1 kernels,
one CTA per utterance, were A/B-timed with this script before their removal: profiles/lattice_kernels_r2_v5.jsonl.)  Algorithmic bytes: SURVEY 8(d)
sum_utt (4 T N + 48 A_lat + 16 S_lat)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pykaldi2_b200 import graphs, synth
from pykaldi2_b200.ops import ops
dev = torch.device('cuda', 0)
rng = np.random.default_rng(3)
N = 5768
Ts = [int(t) for t in (sys.argv[1].split(",") if len(sys.argv) > 1 else [1500, 1200, 900, 600])]
lats, alis = [], []
for T in Ts:
    ali = rng.integers(1, 2 * N, T).astype(np.int32)
    lat, tid2pdf, _ = synth.make_lattice(T, N, rng, num_ali=ali)
    lats.append(graphs.Lattice(lat)); alis.append(ali)
tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 3 + 1, 0)
pred = torch.randn(len(Ts), max(Ts), N, device=dev)
lb = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev, mpe=("smbr", tid2phone, [1]))
alg = sum(4 * T * N for T in Ts) + 48 * lb.total_arcs + 16 * lb.total_states
peak = 6650.0
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
for name, fn in (("lattice MMI", lambda: ops.lattice_mmi(pred, lb)), ("lattice sMBR", lambda: ops.lattice_mpe(pred, lb))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"kernel": name, "impl": "v1", "frames": Ts,
                      "arcs": lb.total_arcs, "states": lb.total_states, "ms_per_call": round(ms, 4),
                      "algorithmic_bytes": alg, "alg_GBps": round(alg / ms / 1e6, 1),
                      "frac_hbm": round(alg / ms / 1e6 / peak, 4)}), flush=True)
