"""ctypes wrapper of the oracle's C restatement (oracle/c/chain_fb.c).  TEST / BASELINE ONLY."""
import ctypes as C

import numpy as np

from . import build_c, chain_ref, lattice_ref

_lib = None
vp, ci, cf = C.c_void_p, C.c_int, C.c_float


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_c.build(verbose=False))
        L.pk2o_den_fb.restype = C.c_double
        L.pk2o_den_fb.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, vp, cf, cf, vp]
        L.pk2o_num_fb.restype = C.c_double
        L.pk2o_num_fb.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, cf, vp]
        L.pk2o_lat_mmi.restype = C.c_double
        L.pk2o_lat_mmi.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, cf, cf, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(vp)


def chain_objf_and_deriv(ll, den, sup_fst, leaky=1e-4, xent_regularize=0.0, weight=1.0):
    """Same contract as chain_ref.chain_objf_and_deriv, computed by the C port (float den, double num)."""
    ll = np.ascontiguousarray(ll, np.float32)
    T, N = ll.shape
    S = den["num_states"]
    g_den = np.zeros((T, N), np.float32)
    arrs = [np.ascontiguousarray(den[k]) for k in ("fwd_off", "fwd_prob", "fwd_pdf", "fwd_state", "initial_probs")]
    z_den = lib().pk2o_den_fb(S, N, T, *[_p(a) for a in arrs], _p(ll), leaky, 1.0, _p(g_den))
    times = np.ascontiguousarray(chain_ref.fst_state_times(sup_fst), np.int32)
    f = {k: np.ascontiguousarray(sup_fst[k], np.int32) for k in ("src", "dst")}
    pdf = np.ascontiguousarray(np.asarray(sup_fst["ilabel"], np.int32) - 1)
    w = np.ascontiguousarray(sup_fst["weight"], np.float32)
    fin = np.ascontiguousarray(sup_fst["final"], np.float32)
    g_num = np.zeros((T, N), np.float32)
    z_num = lib().pk2o_num_fb(int(sup_fst["num_states"]), len(pdf), N, T, int(sup_fst["start"]), _p(f["src"]),
                              _p(f["dst"]), _p(pdf), _p(w), _p(fin), _p(times), _p(ll), 1.0, _p(g_num))
    objf = weight * (z_num - z_den)
    grad = weight * (g_num - g_den) + xent_regularize * weight * g_num
    return float(objf), grad, weight * g_num


def lattice_mmi(ll, lat, tid2pdf, num_ali, lm=1.0, ac=0.2):
    ll = np.ascontiguousarray(ll, np.float32)
    T, N = ll.shape
    times = np.ascontiguousarray(lattice_ref.lattice_state_times(lat), np.int32)
    src, dst, tid = (np.ascontiguousarray(lat[k], np.int32) for k in ("src", "dst", "tid"))
    gc = np.ascontiguousarray(lat["graph_cost"], np.float32)
    fin = np.ascontiguousarray(lat["final_cost"], np.float32)
    t2p = np.ascontiguousarray(tid2pdf, np.int32)
    ali = np.ascontiguousarray(num_ali, np.int32)
    present = set(zip(times[src][tid != 0].tolist(), tid[tid != 0].tolist()))
    keep = np.array([1 if (t, int(a)) in present else 0 for t, a in enumerate(ali)], np.uint8)
    post = np.zeros((T, N), np.float32)
    tot = lib().pk2o_lat_mmi(int(lat["num_states"]), len(src), N, T, _p(src), _p(dst), _p(tid), _p(gc), _p(fin),
                             _p(times), _p(t2p), _p(ali), _p(keep), _p(ll), lm, ac, _p(post))
    return float(tot), post, keep == 0
