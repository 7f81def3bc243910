#!/usr/bin/env python
"""EXPERIMENT (not measured yet): split the batch of one LF-MMI step into the shorter and
the longer half and run them as two micro-batches on two CUDA streams, so that the denominator forward-backward of
the short half (needs many SMs) overlaps the BLSTM recurrence of the long half (64 -> 32 CTAs, latency-bound) and the
backward recurrence of the short half overlaps the denominator of the long half.  Gradients of the two halves
accumulate into .grad; one optimizer step per 64 utterances, i.e. the same update as the single-batch step up to
summation order.  Prints ms per step of the baseline (pipeline.chain_step) and of the two-stream variant.

Knobs to try with it: PK2_DEN_MAX_CLUSTERS (cap on resident denominator clusters, leaves SMs for the other stream's
persistent GEMM CTAs and 16-CTA LSTM clusters), PK2_DEN_HYBRID=0.
"""
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
sb_all = graphs.SupervisionBatch(sups, device=dev)

order = np.argsort(sub, kind="stable")
halves = [order[:len(order) // 2], order[len(order) // 2:]]          # short half, long half
sb_half = [graphs.SupervisionBatch([sups[i] for i in h], device=dev) for h in halves]
idx_half = [torch.as_tensor(h, device=dev) for h in halves]
t_half = [max((frames[i] - 1) // 3 + 1 for i in h) for h in halves]  # output frames of the longest member
streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]


def step_baseline():
    return pipeline.chain_step(model, opt, None, feat, den, opts, wav, woff, foff, sb_all, epoch=0)


def step_two_streams():
    main = torch.cuda.current_stream(dev)
    x, lens = feat.sequence_batch(wav, woff, foff, factor=3, shift=0)        # [B, Tmax', 80]
    ready = torch.cuda.Event(); ready.record(main)
    xs = []
    for k in (0, 1):
        streams[k].wait_event(ready)
        with torch.cuda.stream(streams[k]):
            xk = x.index_select(0, idx_half[k])[:, :t_half[k]].contiguous()
            x.record_stream(streams[k])
            xs.append(xk)
    preds, losses = [None, None], [None, None]
    with torch.cuda.stream(streams[1]):                                      # long half first: it is the critical path
        preds[1] = model(xs[1])
    with torch.cuda.stream(streams[0]):
        preds[0] = model(xs[0])
        losses[0] = ops.ChainObjtiveFunction.apply_batch(preds[0], den, sb_half[0], opts)
        losses[0].backward()
    with torch.cuda.stream(streams[1]):
        losses[1] = ops.ChainObjtiveFunction.apply_batch(preds[1], den, sb_half[1], opts)
        streams[1].wait_stream(streams[0])                                   # .grad of the short half is complete
        losses[1].backward()
    main.wait_stream(streams[0]); main.wait_stream(streams[1])
    pipeline.finish_step(model, opt, None, 5.0)
    return float(losses[0].item()) + float(losses[1].item())


def timeit(fn, steps=6, warm=3):
    for _ in range(warm):
        v = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        v = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, v


if __name__ == "__main__":
    which = sys.argv[1:] or ["baseline", "two_streams"]
    for w in which:
        ms, v = timeit({"baseline": step_baseline, "two_streams": step_two_streams}[w])
        print(json.dumps({"variant": w, "ms_per_step": round(ms, 3), "objf": v}), flush=True)
