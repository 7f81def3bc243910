#!/usr/bin/env python
"""CUDA-event timing of the phases of one LF-MMI training step (bench.py workload, inputs resident):
features | BLSTM forward | chain loss (den + num) | backward | clip + optimizer."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pykaldi2_b200 import graphs, pipeline, synth
from pykaldi2_b200.models.lstm import LSTMAM
from pykaldi2_b200.ops import ops

dev = torch.device("cuda", 0)
durs, wavs, frames, sub, sup_fsts = bench.make_workload(0, bench.BATCH)
den = graphs.DenominatorGraph(synth.make_den_fst(bench.DEN_STATES, bench.N_PDF, bench.DEN_EXTRA, seed=1234), bench.N_PDF)
opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
sups = [graphs.Supervision(f, t, bench.N_PDF) for f, t in zip(sup_fsts, sub)]
torch.manual_seed(0)
model = LSTMAM(bench.FEAT, bench.N_PDF, bench.HID, bench.LAYERS, 0.0, True).to(dev)
model.train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, amsgrad=True)
feat = pipeline.FeaturePipeline(use_cmn=True)
wav_pinned, woff, foff = feat.ex.pack(wavs)
wav = wav_pinned.to(dev)
sb = graphs.SupervisionBatch(sups, device=dev)
names = ["features", "blstm_forward", "chain_loss", "backward", "clip_optimizer"]
acc = np.zeros(len(names))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for it in range(steps + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    x, lens = feat.sequence_batch(wav, woff, foff, factor=3, shift=0)
    ev[1].record()
    pred = model(x, valid_lengths=sub)
    ev[2].record()
    loss = ops.ChainObjtiveFunction.apply_batch(pred, den, sb, opts)
    ev[3].record()
    loss.backward()
    ev[4].record()
    pipeline.finish_step(model, opt, None, 5.0)
    ev[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
acc /= steps
print(json.dumps({"ms": dict(zip(names, [round(float(v), 3) for v in acc])), "total_ms": round(float(acc.sum()), 3)}))
