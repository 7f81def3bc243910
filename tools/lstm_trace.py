#!/usr/bin/env python
"""Per-phase clock64 trace of the cluster LSTM kernels (steps 64..71 of CTA (0,0,0)); profiling aid."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykaldi2_b200 import _lib
from pykaldi2_b200.models.lstm import LSTMAM
dev = torch.device("cuda", 0)
L = _lib.lib()
buf = torch.zeros(256, dtype=torch.int64, device=dev)
model = LSTMAM(80, 512, 512, 1, 0.0, True).to(dev)
x = torch.randn(64, 200, 80, device=dev)
g = torch.randn(64, 200, 512, device=dev) * 1e-3
model(x).backward(g)
L.pk2_lstm_set_profile_buffer(_lib.ptr(buf))
model(x).backward(g)
torch.cuda.synchronize()
L.pk2_lstm_set_profile_buffer(None)
t = buf.cpu().view(2, 8, 16).numpy()
fn = ["h landed (MMA warp)", "MMAs issued", "MMAs retired (epi)", "tmem ld done", "gx landed", "phase1+bar",
      "phase2+bar", "copies issued", "stores issued"]
print("forward (cycles after the previous step's 'copies issued'):")
for s in range(5, 8):
    base = t[0][s - 1][7]
    print("  step %d:" % (64 + s), "  ".join("%s +%d" % (fn[e], t[0][s][e] - base) for e in range(9)))
print("  period:", [int(t[0][s][7] - t[0][s - 1][7]) for s in range(1, 8)])
bn = ["partials landed", "reduced+released", "A tile ready", "MMAs retired", "drained+staged", "copies issued"]
print("backward, thread 0 of CTA 0 (cycles after the previous step's 'copies issued'):")
for s in range(5, 8):
    base = t[1][s - 1][5]
    print("  step %d:" % (64 + s), "  ".join("%s +%d" % (bn[e], t[1][s][e] - base) for e in range(6)),
          " (sum loop done +%d)" % (t[1][s][6] - base))
print("  period:", [int(t[1][s][5] - t[1][s - 1][5]) for s in range(1, 8)])
