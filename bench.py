#!/usr/bin/env python
"""bench.py -- iRTF of the LF-MMI training hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[3], SURVEY.md section 8d "C4"): BLSTM 3x512 (N=5768) LF-MMI chain
training, batch 64 utterances per GPU of LibriSpeech-shaped synthetic audio (durations
clip(gamma(6, 2.05), 1.5, 30) s), synthetic 8192-state denominator FST (mean out-degree 8),
synthetic time-constrained numerator FSTs, 3x frame subsampling, Adam(amsgrad), clip 5.
A step = waveforms -> log-mel fbank -> CMN -> pad/subsample -> BLSTM fwd -> chain den+num
forward-backward -> BLSTM bwd -> (NCCL grad all-reduce) -> clip -> optimizer step.

  value  : hours of audio per wall-clock hour, inputs resident in HBM, all ranks (weak scaling)
  e2e    : same, through the public API with HOST (pinned) waveforms + supervision index arrays
           copied H2D and the loss read back D2H inside the timed region
  roofline: denominator forward-backward kernels (den_forward+den_backward) vs measured HBM peak, plus
           roofline_smem: the same kernels against the shared-memory gather bound they actually run into
  cpu_baseline / --impl reference: the reference's CPU path (numpy fbank restatement, torch-CPU
           nn.LSTM+Linear = the reference model, restated Kaldi chain FB) on a bounded, representative
           sample (every 16th utterance of the length-sorted batch), all physical host cores.
  N > 1:   the global batch is split across ranks by modelled step time (dist.balanced_shards, the
           sampler the trainers use); the line adds allreduce_ms and the per-rank compute-time spread.

Launch: python bench.py --gpus 1 ;  torchrun --nproc-per-node N bench.py --gpus N
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PDF, HID, LAYERS, FEAT = 5768, 512, 3, 80
DEN_STATES, DEN_EXTRA = 8192, 7
BATCH = 64
FACTOR = 3


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(rank, batch, seed=1234, world=1):
    """Synthetic utterances of this rank.  The GLOBAL batch (batch * world utterances, durations seeded
    independently of world) is split across the ranks by dist.balanced_shards -- the product's sampler
    (data.dataloader.BalancedBatchSampler, used by train_se / train_chain): cost = 56 * longest + sum of frames,
    so the rank that holds the longest utterance gets fewer frames.  (The reference's DistributedSampler shards at
    random; SURVEY.md section 7.3 "var-len DP load balance".)"""
    from pykaldi2_b200 import dist as pkdist
    from pykaldi2_b200 import synth
    from pykaldi2_b200.data import fbank as fb
    rng = np.random.default_rng(seed)
    all_durs = synth.make_durations(batch * world, rng)
    if world > 1:
        shards = pkdist.balanced_shards([fb.num_frames(int(round(d * 16000))) for d in all_durs], world, batch)
        durs = np.sort(all_durs[shards[rank]])[::-1]
    else:
        durs = np.sort(all_durs)[::-1]
    rng = np.random.default_rng(seed + 1000 * (rank + 1))
    wavs = synth.make_waveforms(durs, rng)
    frames = [fb.num_frames(len(w)) for w in wavs]
    sub = [(t - 1) // FACTOR + 1 for t in frames]
    sup_fsts = [synth.make_supervision_fst(t, N_PDF, rng) for t in sub]
    return durs, wavs, frames, sub, sup_fsts


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        # NVML in-process (what nvidia-smi reads) when available: spawning nvidia-smi every 0.2 s from a process with
        # a large address space cost a 60 ms hiccup in the first timed step of one run (profiles/bench_r1_v25.json)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            while not self._halt.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
                self._halt.wait(0.05)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


DEN_TRAFFIC_BYTES = 5588676024      # ncu --set full, profiles/ncu_den_full_r2_v10.md (den kernels unchanged since r1 v24: 5585641608)


# ------------------------------------------------------------------------------ CPU arm ----
def host_cores():
    """Physical cores of the box.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit
    that (the N >= 2 reference arms ran on one core)."""
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
    except Exception:
        n = None
    if not n:
        n = max(1, (os.cpu_count() or 2) // 2)
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    # container CPU quota (cgroup v2 cpu.max / v1 cfs quota): more spinning OpenMP threads than the quota allows are
    # throttled together (seen on a 2-GPU box: 24 threads, iRTF 1.2 instead of 25)
    try:
        quota = None
        if os.path.exists("/sys/fs/cgroup/cpu.max"):
            q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
            if q != "max":
                quota = float(q) / float(per)
        elif os.path.exists("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        if quota:
            n = max(1, min(n, int(quota)))
    except Exception:
        pass
    return int(n)


def pick_threads():
    """Thread count for the CPU arm: the physical / quota core count, unless a short probe (the reference model's
    BLSTM forward on a 4 x 100 batch) runs faster with half or a quarter of it -- hidden CPU quotas and busy
    neighbours turn 'all cores' into a 20x slow-down (spinning OpenMP threads), seen on one box of the pool."""
    import torch.nn as nn
    n = host_cores()
    cands = sorted({max(1, n), max(1, n // 2), max(1, n // 4)}, reverse=True)
    if len(cands) == 1:
        return n
    torch.manual_seed(0)
    lstm = nn.LSTM(FEAT, HID, 1, batch_first=True, bidirectional=True)
    x = torch.randn(4, 100, FEAT)
    best, best_t = n, None
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            lstm(x)                                    # thread pool spin-up
            t0 = time.perf_counter()
            for _ in range(3):
                lstm(x)
            dt = time.perf_counter() - t0
            if best_t is None or dt < 0.8 * best_t:   # fewer threads only if clearly faster
                best, best_t = c, dt
    return best


def cpu_sample(durs, wavs, sup_fsts):
    """The bounded sample both CPU legs time: every 16th utterance of the length-sorted batch (4 of 64: a long, two
    medium and a short one -- the padding ratio of the sample is that of the batch)."""
    order = np.argsort(durs)[::-1][::16]
    return [wavs[i] for i in order], [sup_fsts[i] for i in order], float(sum(len(wavs[i]) for i in order)) / 16000.0


def cpu_reference_sample(wavs, sup_fsts, den_fst, threads):
    """The reference's CPU path on the given utterances as ONE minibatch; returns (audio seconds, wall seconds)."""
    import torch.nn as nn
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_port, chain_ref, fbank_ref
    from pykaldi2_b200.data import mel
    torch.set_num_threads(threads)
    n_utts = len(wavs)
    W = mel.mel80_window()
    oden = chain_ref.den_graph_from_fst(den_fst, N_PDF)          # graph construction is one-off: untimed
    torch.manual_seed(0)
    lstm = nn.LSTM(FEAT, HID, LAYERS, batch_first=True, bidirectional=True)
    lin = nn.Linear(2 * HID, N_PDF)
    params = list(lstm.parameters()) + list(lin.parameters())
    opt = torch.optim.Adam(params, lr=1e-4, amsgrad=True)
    t0 = time.perf_counter()
    feats = []
    for w in wavs:
        f = fbank_ref.cmn(fbank_ref.logfbank(w, W)).astype(np.float32)
        feats.append(f[::FACTOR])
    Tm = max(f.shape[0] for f in feats)
    x = np.zeros((n_utts, Tm, FEAT), np.float32)
    for i, f in enumerate(feats):
        x[i, :f.shape[0]] = f
    pred = lin(lstm(torch.from_numpy(x))[0])
    grad = torch.zeros_like(pred)

    def one(i):                                                   # C forward-backward: releases the GIL
        T = feats[i].shape[0]
        ll = pred[i, :T].detach().numpy()
        objf, g, _ = c_port.chain_objf_and_deriv(ll, oden, sup_fsts[i], leaky=1e-4)
        grad[i, :T] = torch.from_numpy(-g.astype(np.float32))
    with ThreadPoolExecutor(max_workers=max(1, min(n_utts, threads))) as ex:
        list(ex.map(one, range(n_utts)))
    pred.backward(grad)
    torch.nn.utils.clip_grad_norm_(params, 5.0)
    opt.step()
    dt = time.perf_counter() - t0
    audio = sum(len(w) for w in wavs) / 16000.0
    return audio, dt


CPU_SAMPLE_TEXT = ("every 16th utterance of the length-sorted 64-utterance batch (4 utterances, one minibatch): numpy "
                   "fbank+CMN, torch-CPU nn.LSTM+Linear fwd/bwd + Adam (the reference model), restated Kaldi chain den/num "
                   "FB (C -O3, one thread per utterance)")


def run_reference(args, rank):
    """--impl reference: CPU path on the box's host cores; rank 0 only."""
    if rank != 0:
        return
    from pykaldi2_b200 import synth
    durs, wavs, frames, sub, sup_fsts = make_workload(0, BATCH)
    den_fst = synth.make_den_fst(DEN_STATES, N_PDF, DEN_EXTRA, seed=1234)
    wv, sf, audio = cpu_sample(durs, wavs, sup_fsts)
    cores = pick_threads()
    for _ in range(1 if args.warmup > 0 else 0):
        cpu_reference_sample(wv[-1:], sf[-1:], den_fst, cores)
    tot_a = tot_t = 0.0
    for _ in range(args.steps):
        a, t = cpu_reference_sample(wv, sf, den_fst, cores)
        tot_a += a; tot_t += t
    irtf = tot_a / tot_t
    line = {
        "impl": "reference", "metric": "iRTF (hours audio/hour) BLSTM LF-MMI", "value": irtf,
        "unit": "hours audio per hour", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BLSTM 3x512 LF-MMI chain, batch 64 var-len utts/GPU, S=8192 den FST (C4)",
                   "sample": "every 16th utterance of the length-sorted batch per step", "sample_audio_s": audio},
        "cpu_baseline": {"value": irtf, "unit": "hours audio per hour", "cores": cores, "kind": "port",
                         "sample": CPU_SAMPLE_TEXT, "sample_audio_s": audio},
        "e2e": {"value": irtf, "unit": "hours audio per hour", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU arm ----
def arm_watchdog(seconds):
    """A hung collective must not burn the box: after `seconds` every thread's Python stack goes to stderr and the
    process exits with status 3 (the launcher then tears the other ranks down)."""
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(seconds, exit=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emulate-shard", default="", help="R/W: single process, no collectives, the utterances rank R of a "
                    "W-rank run would get (shape check of the larger shards of a multi-GPU run)")
    ap.add_argument("--watchdog", type=int, default=int(os.environ.get("PK2_BENCH_WATCHDOG", "420")),
                    help="seconds after which a stuck run dumps its stacks and exits (0 = off)")
    args = ap.parse_args()
    if args.watchdog > 0 and args.impl != "reference":      # the CPU arm cannot hang on a collective; it may be slow
        arm_watchdog(args.watchdog)

    from pykaldi2_b200 import dist as pkdist
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    rank, world, local = pkdist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    from pykaldi2_b200 import _lib, graphs, pipeline, synth
    from pykaldi2_b200.models.lstm import LSTMAM
    from pykaldi2_b200.ops import ops

    t_start = time.perf_counter()

    def log(msg):                           # progress on stderr (stdout carries the one JSON line)
        print("[bench rank %d +%.1fs] %s" % (rank, time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    L = _lib.lib()
    B = args.batch
    if args.emulate_shard:
        er, ew = (int(v) for v in args.emulate_shard.split("/"))
        durs, wavs, frames, sub, sup_fsts = make_workload(er, B, world=ew)
    else:
        durs, wavs, frames, sub, sup_fsts = make_workload(rank, B, world=world)
    log("workload: %d utterances, %d output frames, longest %d" % (len(wavs), sum(sub), max(sub)))
    den_fst = synth.make_den_fst(DEN_STATES, N_PDF, DEN_EXTRA, seed=1234)
    t_den0 = time.perf_counter()
    den = graphs.DenominatorGraph(den_fst, N_PDF)
    t_den_create = time.perf_counter() - t_den0
    log("denominator graph on the device")
    n_arcs = len(den_fst["src"])
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=0.0)
    sups = [graphs.Supervision(f, t, N_PDF) for f, t in zip(sup_fsts, sub)]

    torch.manual_seed(0)
    model = LSTMAM(FEAT, N_PDF, HID, LAYERS, 0.0, True).to(dev)
    model.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-4, amsgrad=True, fused=True)
    pkdist.broadcast_parameters(model)
    averager = pkdist.GradAverager(list(model.parameters())) if world > 1 else None
    feat = pipeline.FeaturePipeline(use_cmn=True)
    wav_pinned, woff, foff = feat.ex.pack(wavs)
    wav_dev = wav_pinned.to(dev)
    sup_dev = graphs.SupervisionBatch(sups, device=dev)
    audio_s = float(sum(len(w) for w in wavs)) / 16000.0
    h2d = wav_pinned.numel() * 4 + sup_dev.h2d_bytes + (len(woff) * 8 + len(foff) * 4) + 2 * 4 * B * ((max(frames) - 1) // FACTOR + 1)

    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage_next():
        """Host -> device transfer of the NEXT step's inputs (pinned waveforms + supervision index arrays) on a
        copy stream, issued while the current step's backward pass runs: what a prefetching data loader does.
        Every step's inputs are copied inside the timed region; only the overlap is gained."""
        with torch.cuda.stream(copy_stream):
            w = wav_pinned.to(dev, non_blocking=True)
            sb = graphs.SupervisionBatch(sups, device=dev)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["next"] = (w, sb, ev)

    step_events = []
    losses = []                             # PendingValue of every step; read one step late (see drain_losses)

    def drain_losses(keep):
        """Read the losses of all but the newest ``keep`` steps: the device->host read of a step's result happens
        after the NEXT step has been enqueued, so the host stays one step ahead of the GPU."""
        out = []
        while len(losses) > keep:
            out.append(losses.pop(0).value())
        return out

    def step(resident):
        if resident:
            loss, _ = pipeline.chain_step(model, optimizer, averager, feat, den, opts, wav_dev, woff, foff, sup_dev, epoch=0,
                                          events=step_events, sync=False)
            losses.append(loss)
            drain_losses(1)
            return
        if "next" not in staged:
            stage_next()
        w, sb, ev = staged.pop("next")
        torch.cuda.current_stream(dev).wait_event(ev)
        w.record_stream(torch.cuda.current_stream(dev))
        sb._keep[0].record_stream(torch.cuda.current_stream(dev))
        loss, _ = pipeline.chain_step(model, optimizer, averager, feat, den, opts, w, woff, foff, sb, epoch=0,
                                      after_backward=stage_next, sync=False)
        losses.append(loss)
        drain_losses(1)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(resident, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step(resident)
        last = drain_losses(0)              # every step's loss is read inside the timed region
        e1.record()
        barrier()
        assert all(np.isfinite(v) for v in last), last
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                      # NVML initialisation happens during the warm-up, not in the timed region
    t_first = None
    for i in range(max(args.warmup, 3)):
        t_s0 = time.perf_counter()
        step(True)
        if i == 0:
            torch.cuda.synchronize()
            t_first = time.perf_counter() - t_s0
            log("first step done (graph tables built, kernels loaded)")
    step(False)
    drain_losses(0)
    torch.cuda.synchronize()
    log("warm-up done")
    if sampler:
        del sampler.rows[:]                  # keep only the samples taken during the timed regions
    t_setup = time.perf_counter() - t_start
    ops.DEN_TIMERS = []
    del step_events[:]
    if averager is not None:
        averager.timers = []
    n0 = L.pk2_launch_count()
    t_res = timed(True, args.steps)
    n1 = L.pk2_launch_count()
    den_ms = [a.elapsed_time(b) for a, b in ops.DEN_TIMERS]
    ops.DEN_TIMERS = None
    compute_ms = float(np.mean([a.elapsed_time(b) for a, b in step_events])) if step_events else float("nan")
    ar_ms = None
    if averager is not None:
        ar_ms = float(np.mean([a.elapsed_time(b) for a, b in averager.timers]))
        averager.timers = None
    # per-rank compute time (start of the step -> backward pass done, before the all-reduce joins the ranks)
    spread = torch.tensor([compute_ms, float(sum(sub)), float(max(sub))], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(spread) for _ in range(world)]
        torch.distributed.all_gather(allv, spread)
        spread_rows = [[float(v) for v in t.tolist()] for t in allv]
        ar_t = torch.tensor([ar_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(ar_t, op=torch.distributed.ReduceOp.MIN)   # the last rank to arrive waits least
        ar_ms = float(ar_t.item())
    else:
        spread_rows = [[float(v) for v in spread.tolist()]]
    log("resident region done: %.2f ms/step" % (1e3 * t_res / args.steps))
    t_e2e = timed(False, args.steps)
    log("e2e region done: %.2f ms/step" % (1e3 * t_e2e / args.steps))
    clocks = sampler.stop() if sampler else None

    tot_audio = torch.tensor([audio_s], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(tot_audio)
    tot_audio = float(tot_audio.item())
    value = tot_audio * args.steps / t_res
    e2e = tot_audio * args.steps / t_e2e

    if rank == 0:
        peak, peak_src = load_peaks()
        S = DEN_STATES
        alg_bytes = sum(sub) * (8 * N_PDF + 8 * S) + 24 * n_arcs
        den_s = float(np.mean(den_ms)) * 1e-3 if den_ms else float("nan")
        achieved = alg_bytes / den_s / 1e9
        line = {
            "metric": "iRTF (hours audio/hour) BLSTM LF-MMI", "value": value, "unit": "hours audio per hour",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BLSTM 3x512 LF-MMI chain, batch 64 var-len utts/GPU, S=8192 den FST (C4)",
                       "global_batch": B * world, "audio_s_per_step": tot_audio, "parallelism": "dp%d" % world,
                       "sharding": "global batch split across ranks by modelled step time (dist.balanced_shards: 56 x longest + sum of frames)",
                       "l2": "no flush: every step streams > 5 GB of activations/workspace, far larger than the 126 MB L2",
                       "optimizer": "Adam(amsgrad) lr 1e-4, clip 5"},
            "e2e": {"value": e2e, "unit": "hours audio per hour", "ms_per_step": 1e3 * t_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "note": "loss of step k read from pinned memory after step k+1 is enqueued (host one step ahead)"},
            "gpu_launches": int(n1 - n0),
            "roofline": {"kernel": "pk2_denfb: den_exp + den_forward_reg2 + den_backward_reg (15 clusters of 8) || den_fb1 (single CTAs)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "ms_per_launch_pair": den_s * 1e3,
                         "algorithmic_bytes": int(alg_bytes),
                         # dram__bytes_read.sum + dram__bytes_write.sum of the four kernels of one pk2_denfb call at B = 64,
                         # one `ncu --set full` capture (profiles/ncu_den_full_r1_v24.md); only valid for the default batch
                         "traffic": DEN_TRAFFIC_BYTES if (B == BATCH and world == 1) else None,
                         "note": "binding bound is shared-memory gather bandwidth + the DSMEM row exchange inside a chain of "
                                 "dependent frames, not HBM (DESIGN.md section 4.4)"},
            # the bound the denominator kernels actually run into: per frame and sequence every
            # arc costs two 4-byte shared-memory gathers in each of the three passes (alpha, beta, gamma); an SM
            # serves 32 conflict-free 4-byte lanes per clock
            "roofline_smem": {"kernel": "pk2_denfb", "bound": "shared-memory gather",
                              "achieved": sum(sub) * n_arcs * 6 / den_s / 1e9, "unit": "G gathers/s",
                              "peak": 148 * 32 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e9 if clocks else 148 * 32 * 1.965,
                              "frac": sum(sub) * n_arcs * 6 / den_s / (148 * 32 * ((clocks["sm_mhz"] if clocks and clocks["sm_mhz"] else 1965.0) * 1e6)),
                              "note": "arcs x 2 gathers x 3 passes x frames / (SMs x 32 lanes x clock); conflict-free ideal"},
            "compute_ms_per_rank": [round(r[0], 3) for r in spread_rows],
            "frames_per_rank": [int(r[1]) for r in spread_rows],
            "tmax_per_rank": [int(r[2]) for r in spread_rows],
            "allreduce_ms": ar_ms,
            # one-off costs outside the timed region: pk2_den_graph_create (CSR -> per-row arc lists, upload) and the first
            # step, which builds the SELL tables of the cluster sizes it uses (incl. the bank-conflict hill-climb of the
            # arc order), loads the kernels and creates the TMA descriptors
            "startup_s": {"den_graph_create": round(t_den_create, 3), "first_step": round(t_first, 3),
                          "to_timed_region": round(t_setup, 3)},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1 and not args.emulate_shard:
            # rank 0 at N = 1 only: at N > 1 the other ranks spin in the closing barrier on the same host cores
            try:
                wv, sf, audio = cpu_sample(durs, wavs, sup_fsts)
                cores = pick_threads()
                a, t = cpu_reference_sample(wv, sf, den_fst, cores)
                line["cpu_baseline"] = {"value": a / t, "unit": "hours audio per hour", "cores": cores,
                                        "kind": "port", "sample": CPU_SAMPLE_TEXT, "sample_audio_s": audio}
            except Exception as e:      # the CPU leg must never take the GPU line down
                line["cpu_baseline"] = {"value": None, "error": repr(e)}
        print(json.dumps(line), flush=True)
    if args.watchdog > 0:
        import faulthandler
        faulthandler.cancel_dump_traceback_later()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
