"""Pin oracle/fbank_ref.py and the host-side mel/chunk/collate logic against golden
vectors produced by the reference's own code (oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import fbank_ref
from pykaldi2_b200.data import mel as melmod

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "fbank_golden.npz"))


def test_mel_window_regenerated_bit_exact():
    ref = np.load(os.path.join(G, "mel80_window.npy"))
    ours = melmod.mel80_window()
    assert ours.dtype == np.float32 and ours.shape == (80, 257)
    assert (ours == ref).all()          # data/mel80_window.txt, bit-exact
    assert (melmod.mel_matrix() == fbank_ref.mel_matrix(ref)).all()


@pytest.mark.parametrize("i", [0, 1, 2, 3, 4])
def test_logfbank_matches_reference(gold, i):
    ref_w = np.load(os.path.join(G, "mel80_window.npy"))
    fb = fbank_ref.logfbank(gold["wav%d" % i], ref_w)
    assert fb.shape == gold["fbank%d" % i].shape
    assert fb.shape[0] == fbank_ref.num_frames(gold["wav%d" % i].shape[0])
    # the reference carries fp32/complex64 rounding: 1e-3 rel / 2e-3 abs
    np.testing.assert_allclose(fb, gold["fbank%d" % i], rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(fbank_ref.cmn(fb), gold["cmn%d" % i], rtol=1e-3, atol=2e-3)


def test_mvn_and_chunks(gold):
    mv = fbank_ref.global_mvn(gold["cmn4"], gold["mvn_mean"], gold["mvn_std"])
    np.testing.assert_allclose(mv, gold["mvn4"], rtol=1e-5, atol=1e-5)
    starts = fbank_ref.utt2seg_index(gold["cmn4"].shape[0])
    assert len(starts) == gold["seg4"].shape[0]
    for k, s in enumerate(starts):
        assert (gold["cmn4"][s:s + 80] == gold["seg4"][k]).all()


def test_num_frames_formula():
    for n, T in [(160000, 999), (240000, 1499), (560, 2), (401, 1), (402, 2), (101, 0)]:
        assert fbank_ref.num_frames(n) == T
