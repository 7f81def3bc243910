"""Sequence-discriminative loss ops: the reference's ``ops.ops`` surface on B200 kernels.

Mirrors reference ops/ops.py: ``MMIFunction`` (:41-75) and ``ChainObjtiveFunction``
(:243-280, spelling kept; ``ChainObjectiveFunction`` is an alias).  Same call pattern
(once per utterance inside a Python loop, results summed, bin/train_se.py:237-251,
bin/train_chain.py:259-278), same return convention (0-dim CPU tensor holding the lattice
log-likelihood / chain objf) and the same backward convention (the saved posterior matrix,
sign flipped, ``grad_out`` ignored).  What changed underneath: no D2H/H2D round trip of
the T x N matrix (ops/ops.py:55,64,255,261,269,271), no PyKaldi -- the forward-backward
runs in libpk2.so on the tensor's own device memory and stream.

``apply_batch`` entry points process a whole padded minibatch [B, Tmax, N] in one C-ABI
call; the per-utterance form is the same code with B = 1.  ``ChainObjtiveFunction.apply_batch`` returns its
0-dim loss on the prediction's device and does not synchronise the host (the per-utterance forms keep the
reference's 0-dim CPU tensor).
There is no CPU fallback: CPU tensors raise.
"""
import numpy as np
import torch as th
from torch.autograd import Function

from .. import _lib
from ..graphs import (ChainTrainingOptions, DenominatorGraph, Lattice, LatticeBatch,  # noqa: F401
                      Supervision, SupervisionBatch, SyntheticLatticeProvider, TidPdfMap)


# bench.py sets this to a list to collect (start, end) CUDA events around the denominator kernels
DEN_TIMERS = None
_SIDE = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = th.cuda.Stream(device=dev)
    return _SIDE[key]


# --------------------------------------------------------------------------- chain ----
def chain_objf_and_deriv(prediction, den_graph, sup_batch, chain_opts, cluster=0):
    """prediction: cuda float32 [B, Tmax, N] raw logits.  Returns
    (objf_per_seq: cuda float64 [B], grad: cuda float32 [B, Tmax, N]) where ``grad`` is
    d(-objf)/d(prediction) = -(w*(gamma_num - gamma_den) + xent_regularize*w*gamma_num)
    (ops/ops.py:265-267,275-280), zero on padded frames."""
    _lib.require_cuda(prediction, "prediction")
    assert prediction.dtype == th.float32 and prediction.dim() == 3
    prediction = prediction.contiguous()
    B, Tmax, N = prediction.shape
    if N != den_graph.num_pdfs():
        raise RuntimeError("prediction has %d columns, den graph %d pdfs" % (N, den_graph.num_pdfs()))
    if B != sup_batch.n_seq:
        raise RuntimeError("batch size %d != number of supervisions %d" % (B, sup_batch.n_seq))
    if max(sup_batch.num_frames_host) > Tmax:
        raise RuntimeError("supervision longer than the network output")
    w = sup_batch.weights[0]
    if any(x != w for x in sup_batch.weights):
        raise RuntimeError("all supervisions of a batch must share one weight")
    L = _lib.lib()
    dev = prediction.device
    grad = th.empty_like(prediction)
    wsb = L.pk2_denfb_workspace_bytes(den_graph.handle, B, Tmax)
    ws = th.empty(wsb, dtype=th.uint8, device=dev)
    logz = th.empty(2, B, dtype=th.float64, device=dev)
    nf = sup_batch._dev["num_frames"]
    ready = th.cuda.Event()
    ready.record(th.cuda.current_stream(dev))
    if DEN_TIMERS is not None:
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
    nf_h = np.ascontiguousarray(sup_batch.num_frames_host, np.int32)      # host copy: length-aware schedule
    _lib.check(L.pk2_denfb(den_graph.handle, _lib.ptr(prediction), _lib.ptr(nf), nf_h.ctypes.data_as(_lib.vp),
                           B, Tmax, Tmax,
                           float(chain_opts.leaky_hmm_coefficient), float(w),
                           _lib.ptr(ws), _lib.ptr(grad), _lib.ptr(logz[0]), int(cluster), _lib.stream()),
               "pk2_denfb")
    if DEN_TIMERS is not None:
        e1.record()
        DEN_TIMERS.append((e0, e1))
    # The numerator forward-backward (latency-bound, one warp per sequence, 8 CTAs) runs on a side stream
    # next to the denominator kernels, on SMs the denominator clusters leave free.
    main = th.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_event(ready)
    with th.cuda.stream(side):
        ab = th.empty(2, max(sup_batch.total_states, 1), dtype=th.float64, device=dev)
        arc_post = th.empty(max(sup_batch.total_arcs, 1), dtype=th.float32, device=dev)
        _lib.check(L.pk2_numfb_post(sup_batch.struct, _lib.ptr(prediction), N, Tmax, _lib.ptr(ab[0]), _lib.ptr(ab[1]),
                                    _lib.ptr(arc_post), _lib.ptr(logz[1]), _lib.vp(side.cuda_stream)), "pk2_numfb_post")
        num_done = th.cuda.Event()
        num_done.record(side)
    prediction.record_stream(side)
    logz.record_stream(side)
    arc_post.record_stream(main)
    main.wait_event(num_done)
    scale = -float(w) * (1.0 + float(chain_opts.xent_regularize))
    _lib.check(L.pk2_numfb_scatter(sup_batch.struct, sup_batch.total_states, _lib.ptr(arc_post), N, Tmax, scale,
                                   _lib.ptr(grad), _lib.stream()), "pk2_numfb_scatter")
    # Kaldi's fallback for a non-finite objective (derivs <- 0, objf <- -10 * weight * T), on the device: no host
    # synchronisation between the loss kernels and the backward pass
    _lib.check(L.pk2_chain_guard(_lib.ptr(logz[0]), _lib.ptr(logz[1]), B, Tmax * N, _lib.ptr(grad), _lib.stream()),
               "pk2_chain_guard")
    objf = w * (logz[1] - logz[0])
    objf = th.where(th.isfinite(objf), objf, -10.0 * w * nf.to(th.float64))
    return objf, grad


class _ChainBatch(Function):
    @staticmethod
    def forward(ctx, prediction, den_graph, sup_batch, chain_opts):
        objf, grad = chain_objf_and_deriv(prediction.detach(), den_graph, sup_batch, chain_opts)
        ctx.save_for_backward(grad)
        # 0-dim tensor on the prediction's device (the per-utterance ChainObjtiveFunction.apply keeps the reference's
        # CPU scalar): reading it back here would stall the host between the loss kernels and the backward pass
        return objf.sum().to(th.float32)

    @staticmethod
    def backward(ctx, grad_out):
        grad_input, = ctx.saved_tensors
        return grad_input, None, None, None


class ChainObjtiveFunction(Function):
    """
        Args:
        loglikes: log-likelihoods from the nnet after the forward operation, cuda [T', N]
        den_graph: the denominator graph for chain model training (graphs.DenominatorGraph)
        supervision: graphs.Supervision (numerator FST, frames_per_sequence, weight)
        chain_opts: graphs.ChainTrainingOptions (leaky_hmm_coefficient, xent_regularize)
    """

    @staticmethod
    def forward(ctx, loglikes, den_graph, supervision, chain_opts):
        ll = loglikes.detach()
        _lib.require_cuda(ll, "loglikes")
        if ll.shape[0] != supervision.frames_per_sequence:
            raise RuntimeError("loglikes has %d rows, supervision %d frames" %
                               (ll.shape[0], supervision.frames_per_sequence))
        sb = SupervisionBatch([supervision], device=ll.device)
        objf, grad = chain_objf_and_deriv(ll.unsqueeze(0), den_graph, sb, chain_opts)
        ctx.save_for_backward(grad[0])
        return th.tensor(float(objf[0].item()))

    @staticmethod
    def backward(ctx, grad_out):
        # the saved tensor already carries the flipped sign (ops/ops.py:278); grad_out is
        # ignored exactly as the reference does.
        grad_input, = ctx.saved_tensors
        return grad_input, None, None, None

    @staticmethod
    def apply_batch(prediction, den_graph, supervisions, chain_opts):
        """prediction cuda [B, Tmax, N]; supervisions: list of Supervision or a SupervisionBatch."""
        if not isinstance(supervisions, SupervisionBatch):
            supervisions = SupervisionBatch(list(supervisions), device=prediction.device)
        return _ChainBatch.apply(prediction, den_graph, supervisions, chain_opts)


ChainObjectiveFunction = ChainObjtiveFunction


# ----------------------------------------------------------------------------- MMI ----
def _tid2pdf_of(trans_model):
    t = getattr(trans_model, "tid2pdf", None)
    if t is None:
        n = trans_model.num_transition_ids()
        t = np.array([-1] + [trans_model.transition_id_to_pdf(i) for i in range(1, n + 1)], np.int32)
        try:
            trans_model.tid2pdf = t
        except Exception:
            pass
    return np.asarray(t, np.int32)


def lattice_mmi(prediction, lat_batch, lm_scale=1.0, ac_scale=0.2):
    """prediction: cuda float32 [B, Tmax, N] (log-prior already subtracted).
    Returns (tot: cuda float64 [B] lattice log-likelihoods, grad [B, Tmax, N] = -post_mat)."""
    _lib.require_cuda(prediction, "prediction")
    assert prediction.dtype == th.float32 and prediction.dim() == 3
    prediction = prediction.contiguous()
    B, Tmax, N = prediction.shape
    if B != lat_batch.n_seq:
        raise RuntimeError("batch size %d != number of lattices %d" % (B, lat_batch.n_seq))
    if max(lat_batch.num_frames_host) > Tmax:
        raise RuntimeError("lattice longer than the network output")
    if lat_batch.max_pdf >= N:
        raise RuntimeError("the transition model maps to pdf %d but the prediction has %d columns" % (lat_batch.max_pdf, N))
    dev = prediction.device
    grad = th.empty_like(prediction)
    tot = th.empty(B, dtype=th.float64, device=dev)
    L = _lib.lib()
    ws = th.empty(L.pk2_latfb_workspace_bytes(lat_batch.total_states, lat_batch.total_arcs, 0), dtype=th.uint8, device=dev)
    _lib.check(L.pk2_latfb_mmi(lat_batch.struct, _lib.ptr(prediction), N, Tmax, Tmax,
                               float(lm_scale), float(ac_scale), _lib.ptr(ws), lat_batch.total_states,
                               lat_batch.total_arcs, lat_batch.total_frames, _lib.ptr(grad), _lib.ptr(tot), _lib.stream()),
               "pk2_latfb_mmi")
    return tot, grad


class _MMIBatch(Function):
    @staticmethod
    def forward(ctx, prediction, lat_batch):
        tot, grad = lattice_mmi(prediction.detach(), lat_batch)
        ctx.save_for_backward(grad)
        return th.tensor(float(tot.sum().item()))

    @staticmethod
    def backward(ctx, grad_out):
        grad_input, = ctx.saved_tensors
        return grad_input, None


class MMIFunction(Function):
    """
        Args:
        loglikes: log likelihoods from the nnet by forwarding the input data, cuda [T, N].
                  Note, the log-prior should be substracted.
        asr_decoder: object with .decode(loglikes) -> {"lattice": graphs.Lattice}
        trans_model: hmm transition model (graphs.TidPdfMap or anything with
                     transition_id_to_pdf / num_transition_ids)
        trans_ids:   alignments in the form of hmm transition ids
    """

    @staticmethod
    def forward(ctx, loglikes, asr_decoder, trans_model, trans_ids):
        ll = loglikes.detach()
        _lib.require_cuda(ll, "loglikes")
        lattice = asr_decoder.decode(ll)["lattice"]
        lb = LatticeBatch([lattice], _tid2pdf_of(trans_model), [np.asarray(trans_ids, np.int32)],
                          device=ll.device)
        tot, grad = lattice_mmi(ll.unsqueeze(0), lb)      # scale (1.0, 0.2): ops/ops.py:58
        ctx.save_for_backward(grad[0])
        return th.tensor(float(tot[0].item()))

    @staticmethod
    def backward(ctx, grad_out):
        grad_input, = ctx.saved_tensors
        return grad_input, None, None, None

    @staticmethod
    def apply_batch(prediction, lat_batch):
        return _MMIBatch.apply(prediction, lat_batch)


def lattice_mpe(prediction, lat_batch, lm_scale=1.0, ac_scale=1.0):
    """sMBR / MPFE over a padded minibatch.  prediction: cuda float32 [B, Tmax, N] log-likelihoods (log-prior
    subtracted); lat_batch: LatticeBatch built with ``mpe=(criterion, tid2phone, silence_phones)``.
    Returns (score [B] float64 = expected frame accuracy per utterance (what sMBRFunction.forward returns),
    grad [B, Tmax, N] = -post_mat, the gradient the reference's backward hands to autograd
    (ops/ops.py:149-156), tot_like [B]).  Scales default to 1: the reference calls no lattice_scale here."""
    _lib.require_cuda(prediction, "prediction")
    assert prediction.dtype == th.float32 and prediction.dim() == 3
    if lat_batch.acc_in is None:
        raise RuntimeError("LatticeBatch was built without mpe=(criterion, tid2phone, silence_phones)")
    prediction = prediction.contiguous()
    B, Tmax, N = prediction.shape
    if B != lat_batch.n_seq:
        raise RuntimeError("batch size %d != number of lattices %d" % (B, lat_batch.n_seq))
    if max(lat_batch.num_frames_host) > Tmax:
        raise RuntimeError("lattice longer than the network output")
    if lat_batch.max_pdf >= N:
        raise RuntimeError("the transition model maps to pdf %d but the prediction has %d columns" % (lat_batch.max_pdf, N))
    dev = prediction.device
    grad = th.empty_like(prediction)
    out = th.empty(2, B, dtype=th.float64, device=dev)
    L = _lib.lib()
    ws = th.empty(L.pk2_latfb_workspace_bytes(lat_batch.total_states, lat_batch.total_arcs, 1), dtype=th.uint8, device=dev)
    _lib.check(L.pk2_latfb_mpe(lat_batch.struct, _lib.ptr(lat_batch.acc_in), _lib.ptr(lat_batch.acc_out),
                               _lib.ptr(prediction), N, Tmax, Tmax, float(lm_scale), float(ac_scale),
                               _lib.ptr(ws), lat_batch.total_states, lat_batch.total_arcs, -1.0, _lib.ptr(grad),
                               _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.stream()), "pk2_latfb_mpe")
    return out[1], grad, out[0]


def _tid2phone_of(trans_model):
    t = getattr(trans_model, "tid2phone", None)
    if t is None:
        n = trans_model.num_transition_ids()
        t = np.zeros(n + 1, np.int32)
        for i in range(1, n + 1):
            t[i] = trans_model.transition_id_to_phone(i)
    return np.asarray(t, np.int32)


class _MPEBatch(Function):
    @staticmethod
    def forward(ctx, prediction, lat_batch):
        score, grad, _ = lattice_mpe(prediction.detach(), lat_batch)
        ctx.save_for_backward(grad)
        return th.tensor(float(score.sum().item()))

    @staticmethod
    def backward(ctx, grad_out):
        grad_input, = ctx.saved_tensors
        return grad_input, None


class sMBRFunction(Function):
    """
        Args:
        loglikes: log likelihoods from the nnet by forwarding the input data, cuda [T, N].
                  Note, the log-prior should be substracted.
        asr_decoder: object with .decode(loglikes) -> {"lattice": graphs.Lattice}
        trans_model: hmm transition model (graphs.TidPdfMap with tid2phone, or anything with
                     transition_id_to_pdf / transition_id_to_phone / num_transition_ids)
        trans_ids:   alignments in the form of hmm transition ids
        criterion: "smbr" or "mpfe"
        silence_phones: slience phone indexes, in the form of list of int

    Reference ops/ops.py:119-156: returns the expected frame accuracy (0-dim CPU tensor); backward hands
    -post_mat to autograd ("flip the sign to maximize the frame accuracy"), grad_out ignored.
    """

    @staticmethod
    def forward(ctx, loglikes, asr_decoder, trans_model, trans_ids, criterion, silence_phones):
        ll = loglikes.detach()
        _lib.require_cuda(ll, "loglikes")
        lattice = asr_decoder.decode(ll)["lattice"]
        lb = LatticeBatch([lattice], _tid2pdf_of(trans_model), [np.asarray(trans_ids, np.int32)], device=ll.device,
                          mpe=(criterion, _tid2phone_of(trans_model), list(silence_phones)))
        score, grad, _ = lattice_mpe(ll.unsqueeze(0), lb)       # no lattice_scale on this path: ops/ops.py:133-143
        ctx.save_for_backward(grad[0])
        return th.tensor(float(score[0].item()))

    @staticmethod
    def backward(ctx, grad_out):
        grad_input, = ctx.saved_tensors
        return grad_input, None, None, None, None, None

    @staticmethod
    def apply_batch(prediction, lat_batch):
        return _MPEBatch.apply(prediction, lat_batch)
