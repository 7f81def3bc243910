#!/usr/bin/env python
"""Kernel timeline of ONE LF-MMI training step (bench.py workload) from CUPTI through torch.profiler: name, stream,
start (us, relative to the first kernel of the step) and duration of every kernel, in launch order of their start
time.  Shows what overlaps what across the streams -- which `ncu` (serialising) cannot.
  python tools/timeline.py > profiles/timeline.csv"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pykaldi2_b200 import graphs, pipeline, synth
from pykaldi2_b200.models.lstm import LSTMAM

dev = torch.device("cuda", 0)
durs, wavs, frames, sub, sup_fsts = bench.make_workload(0, bench.BATCH)
den = graphs.DenominatorGraph(synth.make_den_fst(bench.DEN_STATES, bench.N_PDF, bench.DEN_EXTRA, seed=1234), bench.N_PDF)
opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
sups = [graphs.Supervision(f, t, bench.N_PDF) for f, t in zip(sup_fsts, sub)]
torch.manual_seed(0)
model = LSTMAM(bench.FEAT, bench.N_PDF, bench.HID, bench.LAYERS, 0.0, True).to(dev)
model.train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, amsgrad=True, fused=True)
feat = pipeline.FeaturePipeline(use_cmn=True)
wav_pinned, woff, foff = feat.ex.pack(wavs)
wav = wav_pinned.to(dev)
sb = graphs.SupervisionBatch(sups, device=dev)


pending = []


def step():
    """the bench's loop body: loss read one step late, host one step ahead of the GPU"""
    loss, _ = pipeline.chain_step(model, opt, None, feat, den, opts, wav, woff, foff, sb, epoch=0, sync=False)
    pending.append(loss)
    while len(pending) > 1:
        pending.pop(0).value()


for _ in range(4):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):                       # three steps back to back: the middle one shows the steady state
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
print("start_us,dur_us,end_us,stream,name")
for e in evs:
    st = e.time_range.start - t0
    print("%.1f,%.1f,%.1f,%s,%s" % (st, e.time_range.elapsed_us(), st + e.time_range.elapsed_us(),
                                    getattr(e, "stream", getattr(e, "device_index", "")), e.name.replace(",", ";")[:90]))
