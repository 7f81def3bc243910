#!/usr/bin/env python
"""BASELINE config 3 end to end on one GPU: BLSTM 3x512 lattice-MMI (the bin/train_se.py step; the configuration
the reference's README publishes, 16.7 iRTF on a V100, README.md:48), batch 4 variable-length utterances, synthetic
decoding lattices (K_t ~ U{32..96} states per frame, ~250 arcs per frame, 5 % of the frames without the alignment
arc), CE regulariser 0.1, SGD, clip 5.

A step = waveforms -> fbank -> CMN -> pad -> BLSTM forward -> log-prior -> CE + lattice forward-backward (one batched
call) -> BLSTM backward -> clip -> SGD.  `value`: inputs resident on the device; `e2e`: pinned host waveforms and the
lattice index arrays copied in inside the timed region, loss read back.  Lattices are fixed per utterance (a real
run gets them from the decoder; decoding is outside the path: SURVEY 8a row a12).
  python tools/bench_c3.py [--steps K] [--warmup W] [--criterion mmi|smbr]
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykaldi2_b200 import graphs, pipeline, synth
from pykaldi2_b200.data import fbank as fb
from pykaldi2_b200.models.lstm import LSTMAM
from pykaldi2_b200.ops import ops

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--criterion", default="mmi", choices=["mmi", "smbr"])
args = ap.parse_args()

N, B = 5768, args.batch
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1234)
durs = np.sort(synth.make_durations(B, rng))[::-1]
wavs = synth.make_waveforms(durs, rng)
frames = [fb.num_frames(len(w)) for w in wavs]
lats, alis = [], []
for T in frames:
    ali = rng.integers(1, 2 * N + 1, T).astype(np.int32)
    lat, tid2pdf, _ = synth.make_lattice(T, N, rng, num_ali=ali)
    lats.append(graphs.Lattice(lat)); alis.append(ali)
tid2phone = np.where(np.asarray(tid2pdf) >= 0, np.asarray(tid2pdf) // 3 + 1, 0)
mpe = None if args.criterion == "mmi" else ("smbr", tid2phone, [1])
labels = np.full((B, max(frames)), -100, np.int64)
for b, a in enumerate(alis):
    labels[b, :len(a)] = tid2pdf[a]
y = torch.from_numpy(labels).to(dev)
log_prior = torch.from_numpy(synth.make_log_prior(N, rng)).to(dev)
torch.manual_seed(0)
model = LSTMAM(80, N, 512, 3, 0.0, True).to(dev)
model.train()
opt = torch.optim.SGD(model.parameters(), lr=1e-5, momentum=0.0)
feat = pipeline.FeaturePipeline(use_cmn=True)
wav_pinned, woff, foff = feat.ex.pack(wavs)
wav_dev = wav_pinned.to(dev)
lb_dev = graphs.LatticeBatch(lats, tid2pdf, alis, device=dev, mpe=mpe)
audio_s = float(sum(len(w) for w in wavs)) / 16000.0
LAT_TIMERS = []


def step(resident):
    w = wav_dev if resident else wav_pinned.to(dev, non_blocking=True)
    lb = lb_dev if resident else graphs.LatticeBatch(lats, tid2pdf, alis, device=dev, mpe=mpe)
    x, num_frs = feat.sequence_batch(w, woff, foff)
    pred = model(x, valid_lengths=num_frs)
    ce = pipeline.ce_loss(pred.view(-1, N), y.view(-1), reduction="sum")
    loglikes = pred - log_prior
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    se = ops.MMIFunction.apply_batch(loglikes, lb) if mpe is None else ops.sMBRFunction.apply_batch(loglikes, lb)
    e1.record()
    LAT_TIMERS.append((e0, e1))
    loss = se.cuda() + 0.1 * ce
    loss.backward()
    pipeline.finish_step(model, opt, None, 5.0)
    return float(loss.item())


def timed(resident, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(resident)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


for _ in range(max(args.warmup, 3)):
    step(True)
step(False)
del LAT_TIMERS[:]
t_res = timed(True, args.steps)
lat_ms = float(np.mean([a.elapsed_time(b) for a, b in LAT_TIMERS]))
t_e2e = timed(False, args.steps)
peak = 6650.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
alg = sum(4 * T * N for T in frames) + 48 * lb_dev.total_arcs + 16 * lb_dev.total_states
print(json.dumps({
    "metric": "iRTF (hours audio/hour) BLSTM lattice-" + args.criterion.upper(), "value": audio_s * args.steps / t_res,
    "unit": "hours audio per hour", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
    "ms_per_step": 1e3 * t_res / args.steps, "dtype": "bf16", "data": "synthetic",
    "config": {"workload": "BLSTM 3x512 lattice-%s (train_se.py step), batch %d var-len utts, synthetic lattices (C3)" % (args.criterion, B),
               "frames": frames, "lattice_arcs": lb_dev.total_arcs, "lattice_states": lb_dev.total_states, "audio_s_per_step": audio_s},
    "e2e": {"value": audio_s * args.steps / t_e2e, "ms_per_step": 1e3 * t_e2e / args.steps,
            "h2d_bytes_per_step": int(wav_pinned.numel() * 4 + lb_dev.h2d_bytes), "d2h_bytes_per_step": 16},
    "roofline": {"kernel": "pk2_latfb (arc scores + alpha/beta chains + posterior scatter, incl. the gradient zero fill and "
                           "the host-side launch of the call)", "bound": "hbm",
                 "achieved": alg / (lat_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (lat_ms * 1e-3) / 1e9 / peak,
                 "ms_per_call": lat_ms, "algorithmic_bytes": int(alg)},
    "published_reference": "README.md:48 of the reference: 16.7 iRTF, BLSTM MMI on one V100 (other hardware, real lattices)"}))
