"""BLSTM acoustic model: the reference's ``models.lstm.LSTMAM`` on B200 kernels.

Same constructor, same parameter names / shapes / state_dict keys (``lstm.weight_ih_l0`` ...,
``output_layer.weight``) as reference models/lstm.py:33-54, so its checkpoints load unchanged
(bin/train_ce.py:159-165).  ``forward(x[B,T,F]) -> logits[B,T,N]`` implements the intended
``self.output_layer(output)`` (models/lstm.py:59 has a typo that raises NameError).

Underneath, nn.LSTM / nn.Linear are only parameter containers: forward and backward run in
libpk2.so -- tcgen05 GEMMs for the input projections, the output layer and all weight / input
gradients, and the persistent tcgen05 recurrence kernels of csrc/blstm.cu.  Arithmetic: bf16
operands, fp32 accumulation, fp32 cell state, fp32 master weights (the reference is all-fp32 on
cuDNN; parity budget 1e-3 relative on log-posteriors, see tests/test_gpu_blstm.py).
No CPU fallback: CPU inputs raise.

Three things keep the GEMM part of a step short:
  * the bf16 / per-CTA-packed / transposed copies of the weights are cached on the module and rebuilt only when a
    parameter changed (optimizer step, load_state_dict): one packing per step however many times forward runs;
  * ``forward(data, valid_lengths=...)``: the output layer and its two gradient GEMMs run on the valid (unpadded)
    frames only -- the logits of padded frames are never read by the sequence losses and their gradient is exactly
    zero -- compacted by a row gather and scattered back by the GEMM epilogue; padded logit rows are zero;
  * the weight gradients dW = dY^T X read dY and X where they lie (TN form of the tcgen05 GEMM, MN-major
    descriptors): no transposed copies of the activations.
"""
import ctypes as C

import torch as th
import torch.nn as nn
from torch.autograd import Function

from .. import _lib


# Scheduling hooks for multi-stream callers (pipeline.chain_step_overlapped): optional callables invoked on the host
# just BEFORE a launch is enqueued.  "fwd_recurrence_next": the first layer's recurrence kernel of a forward pass
# comes next in the current stream; "bwd_recurrence_next": the top layer's backward recurrence kernel does.  A caller
# records an event there (it completes when the preceding GEMM has finished, i.e. when the recurrence kernel starts)
# so that kernels of another stream which would take all SMs (the denominator clusters) are held back until the
# 16-CTA LSTM clusters, which need whole GPCs, have been placed.
HOOKS = {}


def _hook(name):
    fn = HOOKS.get(name)
    if fn is not None:
        fn()


def _pad8(n):
    return (n + 7) // 8 * 8


TN_GEMM = True          # weight gradients through the TN form (False: transposed copies + NT form, the round-1 path)


def _gemm(a, b, c, bias, M, N, K, lda, ldb, ldc, bf16_out=False, tn=False, row_map=None):
    flags = (2 if bf16_out else 0) | (4 if tn else 0)
    _lib.check(_lib.lib().pk2_gemm_bf16_ex(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(bias), M, N, K,
                                           lda, ldb, ldc, flags, _lib.ptr(row_map), _lib.stream()), "pk2_gemm_bf16_ex")


def _cast(src):
    dst = th.empty(src.shape, dtype=th.bfloat16, device=src.device)
    _lib.check(_lib.lib().pk2_cast_bf16(_lib.ptr(src), _lib.ptr(dst), src.numel(), _lib.stream()), "pk2_cast_bf16")
    return dst


def _transpose(src, R, Cc, lds):
    """src [R, Cc] (fp32 or bf16, row stride lds) -> bf16 [Cc, pad8(R)] (columns >= R are never read)."""
    ldd = _pad8(R)
    dst = th.empty(Cc, ldd, dtype=th.bfloat16, device=src.device)
    _lib.check(_lib.lib().pk2_transpose_bf16(_lib.ptr(src), 1 if src.dtype == th.bfloat16 else 0, _lib.ptr(dst),
                                             R, Cc, lds, ldd, _lib.stream()), "pk2_transpose_bf16")
    return dst, ldd


def _gather_rows(src, rows, R, Cc):
    """bf16 [R, Cc] = src[rows] (src fp32 or bf16 [*, Cc]); rows None = all rows in order (a plain cast)."""
    dst = th.empty(R, Cc, dtype=th.bfloat16, device=src.device)
    _lib.check(_lib.lib().pk2_gather_rows_bf16(_lib.ptr(src), 1 if src.dtype == th.bfloat16 else 0, _lib.ptr(rows),
                                               _lib.ptr(dst), R, Cc, _lib.stream()), "pk2_gather_rows_bf16")
    return dst


class _Packed(object):
    """bf16 operand copies of the weights in the layouts the kernels read.  Built once per parameter version."""

    def __init__(self, w_out, b_out, lstm_params, L, H):
        self.layers = []
        dev = w_out.device
        lib = _lib.lib()
        bf = dict(dtype=th.bfloat16, device=dev)
        for l in range(L):
            ps = [p.contiguous() for p in lstm_params[8 * l:8 * l + 8]]
            if any(p.dtype != th.float32 for p in ps):
                raise RuntimeError("LSTMAM parameters must be float32 (master weights)")
            I = ps[0].shape[1]
            d = {"I": I, "ldk": _pad8(8 * H),
                 "wih_cat": th.empty(8 * H, I, **bf), "bias_cat": th.empty(8 * H, dtype=th.float32, device=dev),
                 "whh_p": th.empty(2 * 4 * H, H, **bf), "whh_t": th.empty(2 * H, 4 * H, **bf),
                 "whh_tp": th.empty(2 * H, 4 * H, **bf)}
            if l > 0:
                d["wih_t"] = th.empty(I, d["ldk"], **bf)
            ptrs = (C.c_void_p * 8)(*[p.data_ptr() for p in ps])
            _lib.check(lib.pk2_lstm_pack_layer(ptrs, H, I, _lib.ptr(d["wih_cat"]), _lib.ptr(d["bias_cat"]),
                                               _lib.ptr(d["whh_p"]), _lib.ptr(d["whh_t"]), _lib.ptr(d["whh_tp"]),
                                               _lib.ptr(d.get("wih_t")), d["ldk"], _lib.stream()), "pk2_lstm_pack_layer")
            self._keep = ps
            self.layers.append(d)
        w = w_out.contiguous()
        self.w_out = _cast(w)                                                       # [N, 2H]
        self.w_out_t, self.ldw = _transpose(w, w.shape[0], 2 * H, 2 * H)            # [2H, pad8(N)]
        self.b_out = b_out.contiguous()
        self.ready = th.cuda.Event()
        self.ready.record(th.cuda.current_stream(w_out.device))


class _BlstmAM(Function):
    """x[B,T,F] fp32 -> logits[B,T,N] fp32 through L bidirectional LSTM layers + Linear."""

    @staticmethod
    def forward(ctx, x, num_layers, hidden, dropout_p, training, packed, valid, w_out, b_out, *lstm_params):
        _lib.require_cuda(x, "x")
        L, H = num_layers, hidden
        B, T, F = x.shape
        M = B * T
        dev = x.device
        lib = _lib.lib()
        if F % 8 != 0:
            raise RuntimeError("feature dim must be a multiple of 8 (got %d)" % F)
        th.cuda.current_stream(dev).wait_event(packed.ready)
        xin = _cast(x.contiguous().view(M, F))
        saved = {"xin": [], "y": [], "gates": [], "cstate": [], "mask": []}
        N = packed.w_out.shape[0]
        logits = th.empty(B, T, N, dtype=th.float32, device=dev)
        pad_done = None
        if valid is not None:
            # padded logit rows are zero; filled on the side stream next to the recurrence kernels (which leave more
            # than half of the SMs idle) instead of in front of the output GEMM
            main, side = th.cuda.current_stream(dev), _side_stream(dev)
            ev = th.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            logits.record_stream(side)
            _lib.check(lib.pk2_zero_pad_rows(_lib.ptr(logits), _lib.ptr(valid[1]), B, T, 4 * N,
                                             _lib.vp(side.cuda_stream)), "pk2_zero_pad_rows")
            pad_done = th.cuda.Event()
            pad_done.record(side)
        for l in range(L):
            pk = packed.layers[l]
            I = pk["I"]
            gx = th.empty(T, 2, H // 32, B, 128, dtype=th.float32, device=dev)
            _lib.check(lib.pk2_lstm_input_proj(_lib.ptr(xin), _lib.ptr(pk["wih_cat"]), _lib.ptr(pk["bias_cat"]), _lib.ptr(gx),
                                               B, T, I, H, I, _lib.stream()), "pk2_lstm_input_proj")
            y = th.empty(B, T, 2 * H, dtype=th.bfloat16, device=dev)
            gates = th.empty(2, T, B, 4, H, dtype=th.bfloat16, device=dev)
            cstate = th.empty(2, T, B, H, dtype=th.float32, device=dev)
            sync = th.empty(2 * ((B + 31) // 32), dtype=th.int32, device=dev)
            a = _lib.LstmFwdArgs(B, T, H, gx.data_ptr(), pk["whh_p"].data_ptr(), y.data_ptr(), gates.data_ptr(),
                                 cstate.data_ptr(), sync.data_ptr())
            if l == 0:
                _hook("fwd_recurrence_next")
            _lib.check(lib.pk2_lstm_layer_fwd(C.byref(a), _lib.stream()), "pk2_lstm_layer_fwd")
            saved["xin"].append(xin); saved["y"].append(y); saved["gates"].append(gates); saved["cstate"].append(cstate)
            if training and dropout_p > 0 and l < L - 1:
                mask = (th.rand(B, T, 2 * H, device=dev) >= dropout_p).to(th.bfloat16) / (1.0 - dropout_p)
                xin = (y * mask).view(M, 2 * H)
            else:
                mask = None
                xin = y.view(M, 2 * H)
            saved["mask"].append(mask)
            del gx
        if valid is None:
            top, Mv, rows = xin, M, None
        else:
            rows, lens_dev, Mv = valid
            top = _gather_rows(xin, rows, Mv, 2 * H)                    # valid frames only, compact
        if pad_done is not None:
            th.cuda.current_stream(dev).wait_event(pad_done)
        _gemm(top, packed.w_out, logits, packed.b_out, Mv, N, 2 * H, 2 * H, 2 * H, N, row_map=rows)
        ctx.saved = saved
        ctx.top = top
        ctx.rows = rows
        ctx.packed = packed
        ctx.dims = (B, T, F, L, H, N, Mv)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        """Backward pass.  The recurrence kernels of a layer own 64 of the 148 SMs for ~4 ms; the weight /
        bias gradients of the layer ABOVE (K-long TN GEMMs, needed only by the optimizer) run on a side stream on
        the remaining SMs meanwhile.  The recurrence is always launched first so that its 16-CTA clusters get
        whole GPCs; the side-stream GEMMs are capped to the SMs that are left."""
        B, T, F, L, H, N, Mv = ctx.dims
        M = B * T
        saved, packed, rows, top = ctx.saved, ctx.packed, ctx.rows, ctx.top
        dev = dlogits.device
        lib = _lib.lib()
        if N % 8 != 0:
            raise RuntimeError("output size must be a multiple of 8 for the bf16 backward (got %d)" % N)
        main = th.cuda.current_stream(dev)
        side = _side_stream(dev)
        side_cap = max(32, _num_sms(dev) - 2 * (H // 32) * ((B + 31) // 32))
        results = {}

        def on_side(ready_event, fn):
            side.wait_event(ready_event)
            with th.cuda.stream(side):
                lib.pk2_gemm_set_max_ctas(side_cap)
                try:
                    fn()
                finally:
                    lib.pk2_gemm_set_max_ctas(0)

        def mark_ready():
            ev = th.cuda.Event()
            ev.record(main)
            return ev

        def wgrad(dY, ldy, X, ldx, rows_k, m, n):
            """dW[m, n] = dY[rows_k, m]^T X[rows_k, n] (bf16 operands as they lie in memory)."""
            dW = th.empty(m, n, dtype=th.float32, device=dev)
            if TN_GEMM:
                _gemm(dY, X, dW, None, m, n, rows_k, ldy, ldx, n, tn=True)
            else:
                dY_t, ldm = _transpose(dY, rows_k, m, ldy)
                X_t, _ = _transpose(X, rows_k, n, ldx)
                _gemm(dY_t, X_t, dW, None, m, n, rows_k, ldm, ldm, n)
            return dW

        dl = _gather_rows(dlogits.contiguous().view(M, N), rows, Mv, N)  # [Mv, N] bf16, valid frames only
        if rows is None:
            dy = th.empty(M, 2 * H, dtype=th.float32, device=dev)
        else:
            dy = th.zeros(M, 2 * H, dtype=th.float32, device=dev)        # padded frames: zero gradient
        _gemm(dl, packed.w_out_t, dy, None, Mv, 2 * H, N, N, packed.ldw, 2 * H, row_map=rows)

        def out_layer_grads():
            dl.record_stream(side); top.record_stream(side)
            d_w_out = wgrad(dl, N, top, 2 * H, Mv, N, 2 * H)
            d_b_out = th.empty(N, dtype=th.float32, device=dev)
            _lib.check(lib.pk2_colsum_bf16(_lib.ptr(dl), _lib.ptr(d_b_out), Mv, N, _lib.stream()), "pk2_colsum_bf16")
            results["out"] = (d_w_out, d_b_out)

        pending = (mark_ready(), out_layer_grads)
        grads = [None] * (8 * L)
        for l in range(L - 1, -1, -1):
            pk = packed.layers[l]
            I = pk["I"]
            if saved["mask"][l] is not None:
                dy = dy * saved["mask"][l].view(M, 2 * H).float()
            dgates = th.empty(B, T, 2, 4 * H, dtype=th.bfloat16, device=dev)
            sync = th.empty(2 * ((B + 31) // 32), dtype=th.int32, device=dev)
            a = _lib.LstmBwdArgs(B, T, H, dy.data_ptr(), pk["whh_t"].data_ptr(), saved["gates"][l].data_ptr(),
                                 saved["cstate"][l].data_ptr(), dgates.data_ptr(), sync.data_ptr(), pk["whh_tp"].data_ptr())
            if l == L - 1:
                _hook("bwd_recurrence_next")
            if REC_PRIORITY:
                # EXPERIMENT (PK2_LSTM_PRIORITY=1): the recurrence on a high-priority stream.  It and the side stream's
                # weight-gradient GEMM of the layer above become runnable at the same moment (both wait for the
                # input-gradient GEMM); when the GEMM's persistent CTAs are placed first they sit in every GPC and the 16-CTA
                # LSTM clusters wait ~0.27 ms for whole GPCs (profiles/timeline_r2_v15_runahead.csv, layers 2 and 1).  The
                # priority did not change the step time measurably (27.45 against 27.51 ms): not the default.
                rec = _rec_stream(dev)
                ev_in = th.cuda.Event()
                ev_in.record(main)
                rec.wait_event(ev_in)
                for t_ in (dy, dgates, sync):
                    t_.record_stream(rec)
                _lib.check(lib.pk2_lstm_layer_bwd(C.byref(a), _lib.vp(rec.cuda_stream)), "pk2_lstm_layer_bwd")
                ev_out = th.cuda.Event()
                ev_out.record(rec)
                main.wait_event(ev_out)
            else:
                _lib.check(lib.pk2_lstm_layer_bwd(C.byref(a), _lib.stream()), "pk2_lstm_layer_bwd")
            if pending is not None:               # gradients of the layer above: overlap with this recurrence
                on_side(*pending)
                pending = None
            dg2 = dgates.view(M, 8 * H)
            if l > 0:
                dy_next = th.empty(M, I, dtype=th.float32, device=dev)
                _gemm(dg2, pk["wih_t"], dy_next, None, M, I, 8 * H, 8 * H, pk["ldk"], I)
            xin_l, y_l = saved["xin"][l], saved["y"][l]

            def grads_a(l=l, dg2=dg2, xin_l=xin_l, I=I):
                """input weights and biases of layer l"""
                dg2.record_stream(side); xin_l.record_stream(side)
                d_wih = wgrad(dg2, 8 * H, xin_l, I, M, 8 * H, I)
                d_b = th.empty(8 * H, dtype=th.float32, device=dev)
                _lib.check(lib.pk2_colsum_bf16(_lib.ptr(dg2), _lib.ptr(d_b), M, 8 * H, _lib.stream()), "pk2_colsum_bf16")
                results[(l, "a")] = (d_wih, d_b)

            def grads_b(l=l, dg2=dg2, y_l=y_l):
                """recurrent weights of layer l"""
                dg2.record_stream(side); y_l.record_stream(side)
                d_whh = []
                if TN_GEMM:
                    hp = th.empty(M, 2 * H, dtype=th.bfloat16, device=dev)
                    _lib.check(lib.pk2_lstm_hprev(_lib.ptr(y_l), _lib.ptr(hp), B, T, H, _lib.stream()), "pk2_lstm_hprev")
                    for d in range(2):
                        dW = th.empty(4 * H, H, dtype=th.float32, device=dev)
                        _gemm(_lib.ptr_at(dg2, d * 4 * H), _lib.ptr_at(hp, d * H), dW, None, 4 * H, H, M, 8 * H, 2 * H, H, tn=True)
                        d_whh.append(dW)
                else:
                    dg_t, ldm = _transpose(dg2, M, 8 * H, 8 * H)
                    hp_t = th.empty(2 * H, ldm, dtype=th.bfloat16, device=dev)
                    _lib.check(lib.pk2_lstm_hprev_t(_lib.ptr(y_l), _lib.ptr(hp_t), B, T, H, ldm, _lib.stream()),
                               "pk2_lstm_hprev_t")
                    for d in range(2):
                        dW = th.empty(4 * H, H, dtype=th.float32, device=dev)
                        _gemm(dg_t[d * 4 * H:(d + 1) * 4 * H], hp_t[d * H:(d + 1) * H], dW, None, 4 * H, H, M, ldm, ldm, H)
                        d_whh.append(dW)
                results[(l, "b")] = d_whh

            def layer_grads(a=grads_a, b=grads_b):
                a(); b()

            if l > 0:
                pending = (mark_ready(), layer_grads)
                dy = dy_next
            else:
                # bottom layer: no recurrence left to hide behind; the two independent groups run side by side
                on_side(mark_ready(), grads_a)
                grads_b()
        done = th.cuda.Event()
        done.record(side)
        main.wait_event(done)
        d_w_out, d_b_out = results["out"]
        d_w_out.record_stream(main); d_b_out.record_stream(main)
        for l in range(L):
            (d_wih, d_b), d_whh = results[(l, "a")], results[(l, "b")]
            for t in [d_wih, d_b] + d_whh:
                t.record_stream(main)
            grads[8 * l + 0] = d_wih[:4 * H]; grads[8 * l + 4] = d_wih[4 * H:]
            grads[8 * l + 1] = d_whh[0]; grads[8 * l + 5] = d_whh[1]
            # bias_ih and bias_hh receive the same values but must not share storage: AccumulateGrad keeps the
            # tensor it is handed, and in-place clipping would then scale the shared buffer once per parameter
            grads[8 * l + 2] = d_b[:4 * H]; grads[8 * l + 3] = d_b[:4 * H].clone()
            grads[8 * l + 6] = d_b[4 * H:]; grads[8 * l + 7] = d_b[4 * H:].clone()
        ctx.saved = ctx.top = ctx.packed = ctx.rows = None
        return (None, None, None, None, None, None, None, d_w_out, d_b_out) + tuple(grads)


_SIDE = {}
_REC = {}
_SMS = {}
# experiment switch, off by default: measured on one box 27.45 ms per step with it, 27.51 without (profiles/README_r2.md)
REC_PRIORITY = bool(int(__import__("os").environ.get("PK2_LSTM_PRIORITY", "0")))

# Every optimizer step invalidates the packed operand copies.  In-place updates normally bump Tensor._version, but the
# fused optimizers (torch.optim.Adam(fused=True): one multi-tensor kernel) do NOT -- found in profiles/launches_r2_v10.csv
# (three pack launches in eight steps) -- so the cache key also carries a generation counter advanced by a global
# optimizer post-step hook.
_GENERATION = [0]
_FREEZE = bool(int(__import__("os").environ.get("PK2_PACK_FREEZE", "0")))
_STEP_EVENT = {}                 # device index -> CUDA event recorded right after the latest optimizer step


def _bump_generation(*_args, **_kwargs):
    _GENERATION[0] += 1
    if th.cuda.is_available() and th.cuda.is_initialized():
        ev = th.cuda.Event()
        ev.record()              # on the stream the optimizer ran on
        _STEP_EVENT[th.cuda.current_device()] = ev


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(_bump_generation)
except ImportError:                                  # very old torch: repack on every call
    _GENERATION = None


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = th.cuda.Stream(device=dev)
    return _SIDE[key]


def _rec_stream(dev):
    """High-priority stream for the backward recurrence kernels (see _BlstmAM.backward)."""
    key = (dev.type, dev.index)
    if key not in _REC:
        _REC[key] = th.cuda.Stream(device=dev, priority=-1)
    return _REC[key]


def _num_sms(dev):
    key = (dev.type, dev.index)
    if key not in _SMS:
        _SMS[key] = th.cuda.get_device_properties(dev).multi_processor_count
    return _SMS[key]


class LSTMAM(nn.Module):

    def __init__(self, input_size, output_size, hidden_size, num_layers, dropout, bidirectional):
        super(LSTMAM, self).__init__()

        self.input_size = input_size
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.dropout = dropout
        self.bidirectional = bidirectional
        if not bidirectional:
            raise NotImplementedError("the B200 path implements the bidirectional model every reference "
                                      "trainer instantiates (bin/train_ce.py:116: bidirectional=True)")
        self.output_layer = nn.Linear(hidden_size * 2, output_size)
        # parameter container only (names/shapes/init identical to the reference's nn.LSTM)
        self.lstm = nn.LSTM(input_size=self.input_size,
                            hidden_size=self.hidden_size,
                            num_layers=self.num_layers,
                            batch_first=True,
                            dropout=self.dropout,
                            bidirectional=self.bidirectional)
        self._pack, self._pack_key = None, None
        self._valid_cache = {}

    def _flat_params(self):
        out = []
        for l in range(self.num_layers):
            for sfx in ("", "_reverse"):
                for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    out.append(getattr(self.lstm, "%s_l%d%s" % (name, l, sfx)))
        return out

    def _packed(self, flat):
        """Operand copies of the weights, rebuilt only when a parameter may have changed: an optimizer step anywhere in
        the process (generation counter, see _bump_generation), an in-place update (_version), .to() / load_state_dict
        (storage or version)."""
        ps = [self.output_layer.weight, self.output_layer.bias] + flat
        gen = _GENERATION[0] if _GENERATION is not None else object()
        if _FREEZE and self._pack_key is not None:
            return self._pack                        # A/B timing only (PK2_PACK_FREEZE=1): stale operand copies
        key = (gen,) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._pack_key:
            dev = ps[0].device
            main, side = th.cuda.current_stream(dev), _side_stream(dev)
            # The pack kernels (0.2 ms) run on the side stream.  If only an optimizer step happened since the last pack
            # they wait for THAT step's event, not for what the caller has enqueued since (fbank, CMN, gather of the new
            # step), and overlap it; any other change (load_state_dict, in-place edits) waits for the caller's stream.
            step_ev = _STEP_EVENT.get(dev.index)
            only_step = (self._pack_key is not None and step_ev is not None and key[1:] == self._pack_key[1:])
            if only_step:
                side.wait_event(step_ev)
            else:
                ev = th.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
            with th.no_grad(), th.cuda.stream(side):
                self._pack = _Packed(ps[0].detach(), ps[1].detach(), [p.detach() for p in flat], self.num_layers,
                                     self.hidden_size)
            for t in [v for d in self._pack.layers for v in d.values()] + [self._pack.w_out, self._pack.w_out_t]:
                if th.is_tensor(t):
                    t.record_stream(main)            # allocated on the side stream, read by the caller's
            self._pack_key = key
        return self._pack

    def _valid_rows(self, valid_lengths, B, T, dev):
        lens = tuple(int(v) for v in valid_lengths)
        if len(lens) != B or min(lens) < 0 or max(lens) > T:
            raise RuntimeError("valid_lengths must hold one length in [0, T] per sequence")
        key = (lens, T, dev.index)
        hit = self._valid_cache.get(key)
        if hit is None:
            import numpy as np
            rows = np.concatenate([b * T + np.arange(n, dtype=np.int32) for b, n in enumerate(lens)]).astype(np.int32)
            if len(self._valid_cache) > 64:
                self._valid_cache.clear()
            hit = (th.from_numpy(rows).to(dev), th.tensor(lens, dtype=th.int32, device=dev), int(rows.shape[0]))
            self._valid_cache[key] = hit
        return hit

    def forward(self, data, valid_lengths=None):
        """data [B, T, F] -> logits [B, T, N].  ``valid_lengths`` (optional, one int per sequence, in output frames):
        the caller promises to read only logits[b, :valid_lengths[b]] (the sequence losses do); the output layer is
        then evaluated on those frames only and the other rows of the result are zero.  The recurrent layers always
        run over the padding, as the reference's nn.LSTM on the padded batch does (models/lstm.py:58)."""
        if not data.is_cuda:
            raise RuntimeError("pykaldi2_b200.models.lstm.LSTMAM has no CPU path; move the model and data to CUDA")
        flat = self._flat_params()
        valid = None
        if valid_lengths is not None:
            valid = self._valid_rows(valid_lengths, data.shape[0], data.shape[1], data.device)
            if valid[2] == 0 or valid[2] == data.shape[0] * data.shape[1]:
                valid = None
        return _BlstmAM.apply(data.to(th.float32), self.num_layers, self.hidden_size, float(self.dropout),
                              self.training, self._packed(flat), valid, self.output_layer.weight, self.output_layer.bias,
                              *flat)
