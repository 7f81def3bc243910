#!/usr/bin/env python
"""Per-phase cycle trace of the cluster forward LSTM kernel (steps 64..71 of CTA (0,0,0))."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykaldi2_b200 import _lib
from pykaldi2_b200.models.lstm import LSTMAM
dev = torch.device("cuda", 0)
L = _lib.lib()
buf = torch.zeros(128, dtype=torch.int64, device=dev)
L.pk2_lstm_set_profile_buffer(_lib.ptr(buf))
model = LSTMAM(80, 512, 512, 1, 0.0, True).to(dev)
x = torch.randn(64, 200, 80, device=dev)
with torch.no_grad():
    model(x); model(x)
torch.cuda.synchronize()
L.pk2_lstm_set_profile_buffer(None)
t = buf.cpu().view(8, 16).numpy()
names = ["h landed (MMA warp)", "MMAs issued", "MMAs retired (epi)", "tmem ld done", "gx landed", "phase1+bar",
         "phase2+bar", "copies issued", "stores issued"]
for s in range(1, 8):
    base = t[s - 1][7]          # copies issued in the previous step
    print("step %d:" % (64 + s), "  ".join("%s +%d" % (names[e], t[s][e] - base) for e in range(9)))
print("period (copies issued -> copies issued):", [int(t[s][7] - t[s - 1][7]) for s in range(1, 8)])
