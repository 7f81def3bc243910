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


def _gemm(a, b, c, bias, M, N, K, lda, ldb, ldc, bf16_out=False):
    _lib.check(_lib.lib().pk2_gemm_bf16_nt(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(bias), M, N, K,
                                           lda, ldb, ldc, 2 if bf16_out else 0, _lib.stream()), "pk2_gemm_bf16_nt")


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


class _BlstmAM(Function):
    """x[B,T,F] fp32 -> logits[B,T,N] fp32 through L bidirectional LSTM layers + Linear."""

    @staticmethod
    def forward(ctx, x, num_layers, hidden, dropout_p, training, w_out, b_out, *lstm_params):
        _lib.require_cuda(x, "x")
        L, H = num_layers, hidden
        B, T, F = x.shape
        M = B * T
        dev = x.device
        lib = _lib.lib()
        if F % 8 != 0:
            raise RuntimeError("feature dim must be a multiple of 8 (got %d)" % F)
        xin = _cast(x.contiguous().view(M, F))
        saved = {"xin": [], "y": [], "gates": [], "cstate": [], "mask": [], "whh_t": [], "wih_t": []}
        for l in range(L):
            wih_f, whh_f, bih_f, bhh_f, wih_b, whh_b, bih_b, bhh_b = lstm_params[8 * l:8 * l + 8]
            I = wih_f.shape[1]
            wih_cat = _cast(th.cat([wih_f, wih_b], 0).contiguous())                 # [8H, I]
            bias_cat = th.cat([bih_f + bhh_f, bih_b + bhh_b], 0).contiguous()       # [8H]
            gx = th.empty(T, 2, H // 32, B, 128, dtype=th.float32, device=dev)
            _lib.check(lib.pk2_lstm_input_proj(_lib.ptr(xin), _lib.ptr(wih_cat), _lib.ptr(bias_cat), _lib.ptr(gx),
                                               B, T, I, H, I, _lib.stream()), "pk2_lstm_input_proj")
            # recurrent weights packed per CTA: [dir][cta][gate][32][H]
            whh = th.stack([whh_f, whh_b], 0).view(2, 4, H // 32, 32, H).permute(0, 2, 1, 3, 4).contiguous()
            whh_p = _cast(whh.view(2 * 4 * H, H))
            y = th.empty(B, T, 2 * H, dtype=th.bfloat16, device=dev)
            gates = th.empty(2, T, B, 4, H, dtype=th.bfloat16, device=dev)
            cstate = th.empty(2, T, B, H, dtype=th.float32, device=dev)
            sync = th.empty(2 * ((B + 31) // 32), dtype=th.int32, device=dev)
            a = _lib.LstmFwdArgs(B, T, H, gx.data_ptr(), whh_p.data_ptr(), y.data_ptr(), gates.data_ptr(),
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
        N = w_out.shape[0]
        w_out_bf = _cast(w_out.contiguous())
        logits = th.empty(B, T, N, dtype=th.float32, device=dev)
        _gemm(xin, w_out_bf, logits, b_out.contiguous(), M, N, 2 * H, 2 * H, 2 * H, N)
        ctx.saved = saved
        ctx.top_in = xin
        ctx.dims = (B, T, F, L, H, N)
        ctx.save_for_backward(w_out, *lstm_params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        """Backward pass.  The recurrence kernels of a layer own 64 of the 148 SMs for ~4 ms; the weight /
        bias gradients of the layer ABOVE (transposes + K-long GEMMs, needed only by the optimizer) run on a
        side stream on the remaining SMs meanwhile.  The recurrence is always launched first so that its
        16-CTA clusters get whole GPCs; the side-stream GEMMs are capped to the SMs that are left."""
        B, T, F, L, H, N = ctx.dims
        M = B * T
        w_out, *lstm_params = ctx.saved_tensors
        saved = ctx.saved
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

        dl = _cast(dlogits.contiguous().view(M, N))                      # [M, N] bf16
        w_out_t, ldw = _transpose(w_out.contiguous(), N, 2 * H, 2 * H)   # [2H, pad8(N)]
        dy = th.empty(M, 2 * H, dtype=th.float32, device=dev)
        _gemm(dl, w_out_t, dy, None, M, 2 * H, N, N, ldw, 2 * H)
        top_in = ctx.top_in

        def out_layer_grads():
            dl.record_stream(side); top_in.record_stream(side)
            dl_t, ldm = _transpose(dl, M, N, N)                          # [N, pad8(M)]
            top_t, _ = _transpose(top_in, M, 2 * H, 2 * H)               # [2H, pad8(M)]
            d_w_out = th.empty(N, 2 * H, dtype=th.float32, device=dev)
            _gemm(dl_t, top_t, d_w_out, None, N, 2 * H, M, ldm, ldm, 2 * H)
            d_b_out = th.empty(N, dtype=th.float32, device=dev)
            _lib.check(lib.pk2_colsum_bf16(_lib.ptr(dl), _lib.ptr(d_b_out), M, N, _lib.stream()), "pk2_colsum_bf16")
            results["out"] = (d_w_out, d_b_out)

        pending = (mark_ready(), out_layer_grads)
        grads = [None] * (8 * L)
        for l in range(L - 1, -1, -1):
            wih_f, whh_f, _, _, wih_b, whh_b, _, _ = lstm_params[8 * l:8 * l + 8]
            I = wih_f.shape[1]
            if saved["mask"][l] is not None:
                dy = dy * saved["mask"][l].view(M, 2 * H).float()
            whh_t = _cast(th.cat([whh_f.t(), whh_b.t()], 0).contiguous())        # [2H, 4H]
            # same matrix with the 4H index permuted to cta*128 + gate*32 + unit (cluster/DSMEM kernel)
            whh_tp = _cast(th.stack([whh_f, whh_b], 0).view(2, 4, H // 32, 32, H).permute(0, 2, 1, 3, 4)
                           .reshape(2, 4 * H, H).transpose(1, 2).reshape(2 * H, 4 * H).contiguous())
            dgates = th.empty(B, T, 2, 4 * H, dtype=th.bfloat16, device=dev)
            sync = th.empty(2 * ((B + 31) // 32), dtype=th.int32, device=dev)
            a = _lib.LstmBwdArgs(B, T, H, dy.data_ptr(), whh_t.data_ptr(), saved["gates"][l].data_ptr(),
                                 saved["cstate"][l].data_ptr(), dgates.data_ptr(), sync.data_ptr(), whh_tp.data_ptr())
            if l == L - 1:
                _hook("bwd_recurrence_next")
            _lib.check(lib.pk2_lstm_layer_bwd(C.byref(a), _lib.stream()), "pk2_lstm_layer_bwd")
            if pending is not None:               # gradients of the layer above: overlap with this recurrence
                on_side(*pending)
                pending = None
            dg2 = dgates.view(M, 8 * H)
            if l > 0:
                wih_t, ldk = _transpose(th.cat([wih_f, wih_b], 0).contiguous(), 8 * H, I, I)   # [I, 8H]
                dy_next = th.empty(M, I, dtype=th.float32, device=dev)
                _gemm(dg2, wih_t, dy_next, None, M, I, 8 * H, 8 * H, ldk, I)
            xin_l, y_l = saved["xin"][l], saved["y"][l]

            def layer_grads(l=l, dg2=dg2, xin_l=xin_l, y_l=y_l, I=I):
                dg2.record_stream(side); xin_l.record_stream(side); y_l.record_stream(side)
                dg_t, ldm = _transpose(dg2, M, 8 * H, 8 * H)                          # [8H, pad8(M)]
                x_t, _ = _transpose(xin_l, M, I, I)                                    # [I, pad8(M)]
                d_wih = th.empty(8 * H, I, dtype=th.float32, device=dev)
                _gemm(dg_t, x_t, d_wih, None, 8 * H, I, M, ldm, ldm, I)
                hp_t = th.empty(2 * H, ldm, dtype=th.bfloat16, device=dev)
                _lib.check(lib.pk2_lstm_hprev_t(_lib.ptr(y_l), _lib.ptr(hp_t), B, T, H, ldm, _lib.stream()),
                           "pk2_lstm_hprev_t")
                d_whh = th.empty(2, 4 * H, H, dtype=th.float32, device=dev)
                for d in range(2):
                    _gemm(dg_t[d * 4 * H:(d + 1) * 4 * H], hp_t[d * H:(d + 1) * H], d_whh[d], None, 4 * H, H, M, ldm, ldm, H)
                d_b = th.empty(8 * H, dtype=th.float32, device=dev)
                _lib.check(lib.pk2_colsum_bf16(_lib.ptr(dg2), _lib.ptr(d_b), M, 8 * H, _lib.stream()), "pk2_colsum_bf16")
                results[l] = (d_wih, d_whh, d_b)

            if l > 0:
                pending = (mark_ready(), layer_grads)
                dy = dy_next
            else:
                layer_grads()                     # bottom layer: nothing left to overlap with; stay on the main stream
        done = th.cuda.Event()
        done.record(side)
        main.wait_event(done)
        d_w_out, d_b_out = results["out"]
        d_w_out.record_stream(main); d_b_out.record_stream(main)
        for l in range(L):
            d_wih, d_whh, d_b = results[l]
            for t in (d_wih, d_whh, d_b):
                t.record_stream(main)
            grads[8 * l + 0] = d_wih[:4 * H]; grads[8 * l + 4] = d_wih[4 * H:]
            grads[8 * l + 1] = d_whh[0]; grads[8 * l + 5] = d_whh[1]
            # bias_ih and bias_hh receive the same values but must not share storage: AccumulateGrad keeps the
            # tensor it is handed, and in-place clipping would then scale the shared buffer once per parameter
            grads[8 * l + 2] = d_b[:4 * H]; grads[8 * l + 3] = d_b[:4 * H].clone()
            grads[8 * l + 6] = d_b[4 * H:]; grads[8 * l + 7] = d_b[4 * H:].clone()
        ctx.saved = None
        return (None, None, None, None, None, d_w_out, d_b_out) + tuple(grads)


_SIDE = {}
_SMS = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = th.cuda.Stream(device=dev)
    return _SIDE[key]


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

    def _flat_params(self):
        out = []
        for l in range(self.num_layers):
            for sfx in ("", "_reverse"):
                for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    out.append(getattr(self.lstm, "%s_l%d%s" % (name, l, sfx)))
        return out

    def forward(self, data):
        if not data.is_cuda:
            raise RuntimeError("pykaldi2_b200.models.lstm.LSTMAM has no CPU path; move the model and data to CUDA")
        return _BlstmAM.apply(data.to(th.float32), self.num_layers, self.hidden_size, float(self.dropout),
                              self.training, self.output_layer.weight, self.output_layer.bias,
                              *self._flat_params())
