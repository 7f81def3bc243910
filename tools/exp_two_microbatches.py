#!/usr/bin/env python
"""EXPERIMENT: run one LF-MMI step as two half-batches on two CUDA
streams so that the denominator forward-backward of one half (wants many SMs) runs next to the BLSTM recurrence of
the other half (2 clusters of 16 CTAs, latency-bound).

  baseline            pipeline.chain_step (one batch of 64)
  two_streams         round-1 sketch: halves padded to their OWN longest member (not the same update as the batch of
                      64: the reference's BLSTM also runs over the padding, so the padding length changes the result)
  stag:<full|own>:<clusters>:<reserve>
                      staggered schedule, long half first: fwd A | den A || fwd B | bwd A || den B | bwd B, the
                      denominator limited to <clusters> clusters of 8 and <reserve> SMs kept free; "full" pads both
                      halves to the batch's longest utterance => gradients equal to the single-batch step up to the
                      order of summation (checked and printed)

Prints ms per step and, for every variant, the largest relative gradient difference to the baseline.
"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pykaldi2_b200 import graphs, pipeline, synth
from pykaldi2_b200.models import lstm as lstm_mod
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
params = list(model.parameters())
opt = torch.optim.Adam(params, lr=1e-4, amsgrad=True)
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
v_half = [[sub[i] for i in h] for h in halves]


def grads_baseline():
    x, lens = feat.sequence_batch(wav, woff, foff, factor=3, shift=0)
    loss = ops.ChainObjtiveFunction.apply_batch(model(x, valid_lengths=sub), den, sb_all, opts)
    return loss, torch.autograd.grad(loss, params)


def grads_two_streams():
    main = torch.cuda.current_stream(dev)
    x, lens = feat.sequence_batch(wav, woff, foff, factor=3, shift=0)        # [B, Tmax', 80]
    ready = torch.cuda.Event(); ready.record(main)
    xs, preds, losses, gs = [], [None, None], [None, None], [None, None]
    for k in (0, 1):
        streams[k].wait_event(ready)
        with torch.cuda.stream(streams[k]):
            xs.append(x.index_select(0, idx_half[k])[:, :t_half[k]].contiguous())
            x.record_stream(streams[k])
    with torch.cuda.stream(streams[1]):                                      # long half first: it is the critical path
        preds[1] = model(xs[1], valid_lengths=v_half[1])
    with torch.cuda.stream(streams[0]):
        preds[0] = model(xs[0], valid_lengths=v_half[0])
        losses[0] = ops.ChainObjtiveFunction.apply_batch(preds[0], den, sb_half[0], opts)
        gs[0] = torch.autograd.grad(losses[0], params)
    with torch.cuda.stream(streams[1]):
        losses[1] = ops.ChainObjtiveFunction.apply_batch(preds[1], den, sb_half[1], opts)
        gs[1] = torch.autograd.grad(losses[1], params)
    main.wait_stream(streams[0]); main.wait_stream(streams[1])
    for g in gs[0] + gs[1]:
        g.record_stream(main)
    torch._foreach_add_(list(gs[0]), list(gs[1]))
    return losses[0] + losses[1], gs[0]


def make_staggered(fullpad, clusters, reserve):
    def run():
        main = torch.cuda.current_stream(dev)
        sA, sB = streams
        x, lens = feat.sequence_batch(wav, woff, foff, factor=3, shift=0)
        Tfull = x.shape[1]
        ready = torch.cuda.Event(); ready.record(main)
        A, B = 1, 0                                                           # long half first
        tA = Tfull if fullpad else t_half[A]
        tB = Tfull if fullpad else t_half[B]
        ev = {}

        def mark(name):
            def fn():
                e = torch.cuda.Event(); e.record(torch.cuda.current_stream(dev)); ev[name] = e
            return fn
        den.set_sm_budget(clusters, reserve)
        try:
            sA.wait_event(ready); sB.wait_event(ready)
            x.record_stream(sA); x.record_stream(sB)
            with torch.cuda.stream(sA):
                xA = x.index_select(0, idx_half[A])[:, :tA].contiguous()
                predA = model(xA, valid_lengths=v_half[A])
                mark("fwdA")()
            with torch.cuda.stream(sB):
                xB = x.index_select(0, idx_half[B])[:, :tB].contiguous()
                sB.wait_event(ev["fwdA"])
                lstm_mod.HOOKS["fwd_recurrence_next"] = mark("recB")
                predB = model(xB, valid_lengths=v_half[B])
                lstm_mod.HOOKS.pop("fwd_recurrence_next")
            with torch.cuda.stream(sA):
                sA.wait_event(ev["recB"])                                     # B's LSTM clusters are placed first
                lossA = ops.ChainObjtiveFunction.apply_batch(predA, den, sb_half[A], opts)
                lstm_mod.HOOKS["bwd_recurrence_next"] = mark("brecA")
                gA = torch.autograd.grad(lossA, params)
                lstm_mod.HOOKS.pop("bwd_recurrence_next")
            with torch.cuda.stream(sB):
                sB.wait_event(ev["brecA"])                                    # A's backward clusters are placed first
                lossB = ops.ChainObjtiveFunction.apply_batch(predB, den, sb_half[B], opts)
                gB = torch.autograd.grad(lossB, params)
        finally:
            lstm_mod.HOOKS.clear()
            den.set_sm_budget(0, 0)
        main.wait_stream(sA); main.wait_stream(sB)
        for g in gA + gB:
            g.record_stream(main)
        lossB = lossB.clone(); lossB.record_stream(main)
        torch._foreach_add_(list(gA), list(gB))
        return lossA + lossB, gA
    return run


def full_step(gfn):
    loss, gs = gfn()
    for p, g in zip(params, gs):
        p.grad = g
    pipeline.finish_step(model, opt, None, 5.0)
    return float(loss.item())


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


def variant(name):
    if name == "baseline":
        return grads_baseline
    if name == "two_streams":
        return grads_two_streams
    _, pad, cl, rs = name.split(":")
    return make_staggered(pad == "full", int(cl), int(rs))


if __name__ == "__main__":
    which = sys.argv[1:] or ["baseline", "two_streams", "stag:full:12:32", "stag:full:0:0", "stag:full:10:48", "stag:own:12:32"]
    print(json.dumps({"CUDA_DEVICE_MAX_CONNECTIONS": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS")}), flush=True)
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    _, gref = grads_baseline()
    gref = [g.clone() for g in gref]
    torch.cuda.synchronize()
    for w in which:
        model.load_state_dict(state0)
        gfn = variant(w)
        _, gs = gfn()
        torch.cuda.synchronize()
        diff = max(float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(gs, gref))
        ms, v = timeit(lambda: full_step(gfn))
        print(json.dumps({"variant": w, "ms_per_step": round(ms, 3), "objf": v, "max_rel_grad_diff_vs_baseline": diff}),
              flush=True)
