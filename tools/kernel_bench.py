#!/usr/bin/env python
"""Per-kernel timings on the GPU box (CUDA events, warm, on the launching stream).

  python tools/kernel_bench.py den      # denominator FB: uniform cluster sizes vs the mixed schedule
  python tools/kernel_bench.py lstm     # recurrent kernels: us per step, forward and backward
  python tools/kernel_bench.py gemm     # tcgen05 GEMM TFLOP/s on the model's shapes
  python tools/kernel_bench.py fbank    # fbank GB/s
  python tools/kernel_bench.py cudnn    # the reference model on cuDNN/cuBLAS (fp32 / tf32 / bf16) vs LSTMAM, fwd and fwd+bwd
Prints one JSON object per measurement (copied into profiles/ by hand).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def bench_den():
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    N, S = 5768, 8192
    dev = torch.device("cuda", 0)
    fst = synth.make_den_fst(S, N, 7, seed=1234)
    den = graphs.DenominatorGraph(fst, N)
    A = len(fst["src"])
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4)
    rng = np.random.default_rng(1234)
    durs = synth.make_durations(64, rng)
    Ts = [int((int(d * 100) - 1) // 3 + 1) for d in durs]
    sups = [graphs.Supervision(synth.make_supervision_fst(t, N, rng), t, N) for t in Ts]
    sb = graphs.SupervisionBatch(sups, device=dev)
    pred = torch.randn(64, max(Ts), N, device=dev) * 2
    alg = sum(Ts) * (8 * N + 8 * S) + 24 * A
    for K in (8, 4, 0):
        ms = timeit(lambda: ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=K), iters=3, warm=1)
        # uniform T batch to get the per-frame cost of a cluster size
        print(json.dumps({"kernel": "denfb+numfb", "cluster": K if K else "auto", "ms": ms, "frames": sum(Ts),
                          "max_T": max(Ts), "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / peaks()["hbm_gbs"]}))
    Tu = 200
    supu = [graphs.Supervision(synth.make_supervision_fst(Tu, N, rng), Tu, N) for _ in range(16)]
    sbu = graphs.SupervisionBatch(supu, device=dev)
    predu = torch.randn(16, Tu, N, device=dev) * 2
    for K in (1, 2, 4, 8):
        ms = timeit(lambda: ops.chain_objf_and_deriv(predu, den, sbu, opts, cluster=K), iters=3, warm=1)
        print(json.dumps({"kernel": "denfb uniform T=200 B=16", "cluster": K, "ms": ms, "us_per_frame": 1e3 * ms / Tu}))


def bench_den1():
    """The bench.py batch through the automatic denominator schedule, twice: the target of `ncu --set full -k regex:den_`
    (profiles/ncu_den_full_r2_*.md) -- the capture then holds exactly the kernels of one pk2_denfb call."""
    import bench
    from pykaldi2_b200 import graphs, synth
    from pykaldi2_b200.ops import ops
    dev = torch.device("cuda", 0)
    durs, wavs, frames, sub, sup_fsts = bench.make_workload(0, bench.BATCH)
    den = graphs.DenominatorGraph(synth.make_den_fst(bench.DEN_STATES, bench.N_PDF, bench.DEN_EXTRA, seed=1234), bench.N_PDF)
    opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4)
    sb = graphs.SupervisionBatch([graphs.Supervision(f, t, bench.N_PDF) for f, t in zip(sup_fsts, sub)], device=dev)
    pred = torch.randn(bench.BATCH, max(sub), bench.N_PDF, device=dev) * 2
    ms = timeit(lambda: ops.chain_objf_and_deriv(pred, den, sb, opts, cluster=0), iters=1, warm=1)
    print(json.dumps({"kernel": "pk2_denfb auto schedule, bench batch", "ms": ms, "frames": sum(sub)}))


def bench_lstm():
    from pykaldi2_b200.models.lstm import LSTMAM
    dev = torch.device("cuda", 0)
    for B, T in ((64, 300), (64, 900), (256, 80), (4, 1500)):
        model = LSTMAM(80, 5768, 512, 3, 0.0, True).to(dev)
        x = torch.randn(B, T, 80, device=dev)
        g = torch.randn(B, T, 5768, device=dev) * 1e-3

        def fwd():
            with torch.no_grad():
                model(x)

        def fwdbwd():
            out = model(x)
            out.backward(g)
        mf = timeit(fwd, 3, 1)
        mfb = timeit(fwdbwd, 3, 1)
        flops = 125.5e6 * B * T
        print(json.dumps({"kernel": "LSTMAM 3x512", "B": B, "T": T, "fwd_ms": mf, "fwd_bwd_ms": mfb,
                          "fwd_us_per_step_layer": 1e3 * mf / (3 * T), "train_TFLOPs": flops / mfb / 1e9}))


def bench_cudnn():
    """The bar SURVEY 2.3 sets for row a8: the reference model itself -- torch.nn.LSTM (cuDNN) + nn.Linear (cuBLAS),
    reference models/lstm.py:46-58 with cudnn.benchmark as bin/train_ce.py:118 sets it -- on the same box and shapes,
    fp32 (what the reference runs), fp32 with TF32 allowed, and bf16 parameters; next to this repo's LSTMAM."""
    import torch.nn as nn
    from pykaldi2_b200.models.lstm import LSTMAM
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True

    class Ref(nn.Module):
        def __init__(self):
            super().__init__()
            self.lstm = nn.LSTM(80, 512, 3, batch_first=True, dropout=0.0, bidirectional=True)
            self.output_layer = nn.Linear(1024, 5768)

        def forward(self, x):
            return self.output_layer(self.lstm(x)[0])

    for B, T in ((64, 882), (256, 80), (4, 1500)):
        x = torch.randn(B, T, 80, device=dev)
        g = torch.randn(B, T, 5768, device=dev) * 1e-3
        rows = {}
        for name in ("cudnn_fp32", "cudnn_tf32", "cudnn_bf16", "ours_bf16"):
            torch.backends.cuda.matmul.allow_tf32 = name == "cudnn_tf32"
            torch.backends.cudnn.allow_tf32 = name == "cudnn_tf32"
            torch.manual_seed(0)
            if name == "ours_bf16":
                model, xx, gg = LSTMAM(80, 5768, 512, 3, 0.0, True).to(dev), x, g
            elif name == "cudnn_bf16":
                model, xx, gg = Ref().to(dev).to(torch.bfloat16), x.to(torch.bfloat16), g.to(torch.bfloat16)
            else:
                model, xx, gg = Ref().to(dev), x, g
            model.train()

            def fwd():
                with torch.no_grad():
                    model(xx)

            def fwdbwd():
                model(xx).backward(gg)
                for p in model.parameters():
                    p.grad = None
            rows[name] = (timeit(fwd, 3, 2), timeit(fwdbwd, 3, 2))
            del model
        print(json.dumps({"kernel": "BLSTM 3x512 + Linear 5768", "B": B, "T": T,
                          **{k + "_fwd_ms": round(v[0], 3) for k, v in rows.items()},
                          **{k + "_fwd_bwd_ms": round(v[1], 3) for k, v in rows.items()},
                          "speedup_vs_cudnn_fp32_fwd_bwd": round(rows["cudnn_fp32"][1] / rows["ours_bf16"][1], 2),
                          "speedup_vs_cudnn_bf16_fwd_bwd": round(rows["cudnn_bf16"][1] / rows["ours_bf16"][1], 2)}), flush=True)


def bench_gemm():
    from pykaldi2_b200.models import lstm as L
    dev = torch.device("cuda", 0)
    pk = peaks()
    for M, N, K in ((56448, 4096, 1024), (56448, 5768, 1024), (56448, 1024, 5768), (5768, 1024, 56448), (4096, 1024, 56448), (57600, 4096, 1024), (57600, 5768, 1024), (57600, 1024, 5768), (4096, 1024, 57600),
                    (5768, 1024, 57600), (2048, 512, 57600), (57600, 1024, 4096), (20480, 4096, 1024), (8192, 8192, 8192)):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        b = torch.randn(N, K, device=dev).to(torch.bfloat16)
        c = torch.empty(M, N, device=dev)
        ms = timeit(lambda: L._gemm(a, b, c, None, M, N, K, K, K, N), 5, 2)
        tf = 2.0 * M * N * K / ms / 1e9
        ms_t = timeit(lambda: torch.matmul(a, b.t()), 5, 2)
        print(json.dumps({"kernel": "gemm_bf16_nt", "M": M, "N": N, "K": K, "ms": ms, "TFLOPs": tf,
                          "frac_of_measured_peak": tf / pk["bf16_tflops"], "cublas_ms": ms_t}))


def bench_fbank():
    from pykaldi2_b200 import synth
    from pykaldi2_b200.data import fbank
    rng = np.random.default_rng(0)
    wavs = synth.make_waveforms(synth.make_durations(64, rng), rng)
    ex = fbank.FbankExtractor()
    buf, woff, foff = ex.pack(wavs)
    wd = buf.cuda()
    ms = timeit(lambda: ex.extract(wd, woff, foff), 10, 3)
    byts = 4 * buf.numel() + 320 * int(foff[-1])
    print(json.dumps({"kernel": "fbank", "ms": ms, "audio_s": buf.numel() / 16000.0, "GBps": byts / ms / 1e6,
                      "frac_hbm": byts / ms / 1e6 / peaks()["hbm_gbs"]}))


def bench_ce():
    """CE softmax + NLL forward/backward on BASELINE config 2 (256 chunks x 80 frames, N = 5768)."""
    from pykaldi2_b200 import _lib
    dev = torch.device("cuda", 0)
    R, N = 256 * 80, 5768
    logits = torch.randn(R, N, device=dev)
    labels = torch.randint(0, N, (R,), device=dev)
    loss = torch.empty(R, device=dev)
    grad = torch.empty_like(logits)
    L = _lib.lib()
    ms = timeit(lambda: _lib.check(L.pk2_ce_softmax(_lib.ptr(logits), _lib.ptr(labels), R, N, 1.0, _lib.ptr(loss),
                                                     _lib.ptr(grad), _lib.stream()), "ce"), iters=10, warm=3)
    gbs = 2 * 4 * R * N / ms / 1e6
    print(json.dumps({"kernel": "ce_softmax", "rows": R, "cols": N, "ms": ms, "GBps": gbs, "frac_hbm": gbs / peaks()["hbm_gbs"]}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["den", "lstm", "gemm", "fbank"]
    for w in which:
        {"den": bench_den, "lstm": bench_lstm, "gemm": bench_gemm, "fbank": bench_fbank, "ce": bench_ce,
         "cudnn": bench_cudnn, "den1": bench_den1}[w]()
