"""The oracle's plain-C restatement (oracle/c/chain_fb.c, used as the timed CPU baseline) agrees
with the float64 numpy oracle that the brute-force / autograd tests pin."""
import numpy as np

from oracle import c_port, chain_ref, lattice_ref
from pykaldi2_b200 import synth


def test_c_chain_matches_numpy_oracle():
    rng = np.random.default_rng(4)
    N, T = 60, 28
    den = chain_ref.den_graph_from_fst(synth.make_den_fst(96, N, 5, seed=9), N)
    sup = synth.make_supervision_fst(T, N, rng)
    ll = rng.normal(0, 2, (T, N)).astype(np.float32)
    o1, g1, x1 = chain_ref.chain_objf_and_deriv(ll, den, sup, leaky=1e-4, xent_regularize=0.1)
    o2, g2, x2 = c_port.chain_objf_and_deriv(ll, den, sup, leaky=1e-4, xent_regularize=0.1)
    np.testing.assert_allclose(o2, o1, rtol=1e-4)
    np.testing.assert_allclose(g2, g1, rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(x2, x1, rtol=1e-3, atol=2e-6)


def test_c_lattice_matches_numpy_oracle():
    rng = np.random.default_rng(8)
    N, T = 80, 25
    for eps in (0.0, 0.2):
        lat, t2p, ali = synth.make_lattice(T, N, rng, kmin=5, kmax=10, ali_drop=0.2, eps_frac=eps)
        ll = rng.normal(0, 3, (T, N)).astype(np.float32)
        tot, post, drop, _ = lattice_ref.lattice_fb_mmi(ll, lat, t2p, ali)
        tot2, post2, drop2 = c_port.lattice_mmi(ll, lat, t2p, ali)
        assert (drop == drop2).all()
        np.testing.assert_allclose(tot2, tot, rtol=1e-6)
        np.testing.assert_allclose(post2, post, rtol=1e-3, atol=1e-6)
