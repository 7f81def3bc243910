"""Chain supervision builder: alignment -> phones/durations -> proto-supervision -> numerator FST.
Restated from Kaldi's chain-supervision.cc / hmm-utils.cc (parity unpinned: no Kaldi here); pinned by hand-built
cases whose arrays were derived on paper, and by brute-force enumeration of the accepted pdf sequences."""
import itertools

import numpy as np
import pytest

from pykaldi2_b200 import chain_supervision as cs
from pykaldi2_b200.reader import kaldi_io

CHAIN_TM = """<TransitionModel>
<Topology>
<TopologyEntry>
<ForPhones> 1 2 3 </ForPhones>
<State> 0 <ForwardPdfClass> 0 <SelfLoopPdfClass> 1 <Transition> 0 0.5 <Transition> 1 0.5 </State>
<State> 1 </State>
</TopologyEntry>
</Topology>
<Tuples> 3
1 0 0 1
2 0 2 3
3 0 4 5
</Tuples>
<LogProbs>  [ 0 -0.69 -0.69 -0.69 -0.69 -0.69 -0.69 ] </LogProbs>
</TransitionModel>
"""

ALI_TM = """<TransitionModel>
<Topology>
<TopologyEntry>
<ForPhones> 1 2 </ForPhones>
<State> 0 <PdfClass> 0 <Transition> 0 0.75 <Transition> 1 0.25 </State>
<State> 1 <PdfClass> 1 <Transition> 1 0.75 <Transition> 2 0.25 </State>
<State> 2 <PdfClass> 2 <Transition> 2 0.75 <Transition> 3 0.25 </State>
<State> 3 </State>
</TopologyEntry>
</Topology>
<Triples> 6
1 0 0
1 1 1
1 2 2
2 0 3
2 1 4
2 2 5
</Triples>
<LogProbs>  [ 0 0 0 0 0 0 0 0 0 0 0 0 0 ] </LogProbs>
</TransitionModel>
"""

BIPHONE_TREE = ("ContextDependency 2 1 ToPdf TE 1 4 ( NULL "
                "SE 0 [ 0 1 ] { TE -1 2 ( CE 0 CE 1 ) TE -1 2 ( CE 6 CE 7 ) } "
                "TE -1 2 ( CE 2 CE 3 ) TE -1 2 ( CE 4 CE 5 ) ) EndContextDependency")


def mono_tree():
    return kaldi_io.ContextDependency.monophone({(1, 0): 0, (1, 1): 1, (2, 0): 2, (2, 1): 3, (3, 0): 4, (3, 1): 5})


def accepted(fst, T):
    """All label sequences of length T the FST accepts (brute force over its paths)."""
    out_arcs = {}
    for s, d, l in zip(fst["src"], fst["dst"], fst["ilabel"]):
        out_arcs.setdefault(int(s), []).append((int(d), int(l)))
    res = set()

    def walk(s, seq):
        if len(seq) == T:
            if np.isfinite(fst["final"][s]):
                res.add(tuple(seq))
            return
        for d, l in out_arcs.get(s, []):
            walk(d, seq + [l])
    walk(int(fst["start"]), [])
    return res


def test_transition_model_flags():
    tm = kaldi_io.read_transition_model_text(ALI_TM)
    assert tm["tid2phone"].tolist() == [0] + [1] * 6 + [2] * 6
    assert tm["tid_is_self_loop"].tolist() == [False] + [True, False] * 6
    assert tm["tid_is_final"].tolist() == [False] + [False] * 5 + [True] + [False] * 5 + [True]
    assert tm["tid2state"].tolist() == [0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
    assert tm["pdf_class"][1] == [(0, 0), (1, 1), (2, 2), None]
    ctm = kaldi_io.read_transition_model_text(CHAIN_TM)
    assert ctm["pdf_class"][2] == [(0, 1), None]
    assert ctm["tid2pdf"].tolist() == [-1, 1, 0, 3, 2, 5, 4]
    assert ctm["tid_is_final"].tolist() == [False, False, True, False, True, False, True]


def test_split_to_phones_plain_and_reordered():
    tm = kaldi_io.read_transition_model_text(ALI_TM)
    plain = [1, 1, 2, 3, 4, 5, 5, 6, 8, 10, 11, 12]
    assert cs.split_to_phones(tm, plain) == [(1, 0, 8), (2, 8, 4)]
    reordered = [2, 1, 1, 4, 3, 6, 5, 5, 8, 10, 12, 11]
    assert cs.split_to_phones(tm, reordered) == [(1, 0, 8), (2, 8, 4)]
    # the same phone twice in a row is two segments (the final transition separates them)
    assert cs.split_to_phones(tm, [2, 4, 6, 2, 4, 5, 6]) == [(1, 0, 3), (1, 3, 4)]
    with pytest.raises(ValueError):
        cs.split_to_phones(tm, [1, 99])
    assert cs.split_to_phones(tm, []) == []


def test_alignment_to_proto_supervision_hand_case():
    opts = cs.SupervisionOptions(left_tolerance=5, right_tolerance=5, frame_subsampling_factor=3)
    proto = cs.alignment_to_proto_supervision(opts, [1, 2], [8, 4])
    assert proto.allowed_phones == [[1], [1, 2], [1, 2], [1, 2]]
    assert proto.src.tolist() == [0, 1, 1, 2] and proto.dst.tolist() == [1, 1, 2, 2]
    assert proto.label.tolist() == [1, 1, 2, 2] and proto.num_states == 3 and proto.phones == [1, 2]
    # no tolerance, no subsampling: every frame allows exactly the aligned phone
    exact = cs.alignment_to_proto_supervision(cs.SupervisionOptions(0, 0, 1), [3, 1, 3], [2, 1, 2])
    assert exact.allowed_phones == [[3], [3], [1], [3], [3]]
    # a phone shorter than the subsampling step may get no frame of its own
    short = cs.alignment_to_proto_supervision(cs.SupervisionOptions(0, 0, 3), [1, 2, 3], [4, 1, 4])
    assert short.allowed_phones == [[1], [1], [3]]
    with pytest.raises(ValueError):
        cs.alignment_to_proto_supervision(opts, [1, 2], [3])


def test_supervision_fst_hand_case_bit_exact():
    """phones (1, 2), durations (8, 4), tolerance 5, subsampling 3, monophone tree: arrays derived by hand."""
    opts = cs.SupervisionOptions(5, 5, 3)
    tm = kaldi_io.read_transition_model_text(CHAIN_TM)
    proto = cs.alignment_to_proto_supervision(opts, [1, 2], [8, 4])
    fst = cs.proto_supervision_to_supervision(mono_tree(), tm, proto)
    assert fst["num_states"] == 7 and fst["start"] == 0
    assert fst["src"].tolist() == [0, 1, 1, 2, 3, 3, 4, 5]
    assert fst["dst"].tolist() == [1, 2, 3, 4, 4, 5, 6, 6]
    assert fst["ilabel"].tolist() == [1, 3, 2, 4, 3, 2, 4, 3]
    assert fst["state_times"].tolist() == [0, 1, 2, 2, 3, 3, 4]
    assert np.isfinite(fst["final"]).tolist() == [False] * 6 + [True] and fst["final"][6] == 0
    assert (fst["weight"] == 0).all()
    assert accepted(fst, 4) == {(1, 3, 4, 4), (1, 2, 3, 4), (1, 2, 2, 3)}


@pytest.mark.parametrize("phones,durs,tol,factor", [([1, 2, 3], [5, 3, 7], 2, 3), ([2, 2, 1], [3, 4, 2], 1, 1),
                                                   ([3, 1], [6, 6], 5, 3), ([1, 2, 1, 3], [2, 2, 2, 3], 0, 1)])
def test_supervision_language_by_brute_force(phones, durs, tol, factor):
    """The FST accepts exactly the pdf sequences  fwd_1 sl_1^a1 fwd_2 sl_2^a2 ...  of total length T whose phones lie
    inside the per-frame allowed sets (independent enumeration over the durations a_i)."""
    opts = cs.SupervisionOptions(tol, tol, factor)
    tm = kaldi_io.read_transition_model_text(CHAIN_TM)
    proto = cs.alignment_to_proto_supervision(opts, phones, durs)
    T = len(proto.allowed_phones)
    fst = cs.proto_supervision_to_supervision(mono_tree(), tm, proto)
    pdf = {1: (0, 1), 2: (2, 3), 3: (4, 5)}
    want = set()
    for reps in itertools.product(range(T), repeat=len(phones)):
        if sum(reps) + len(phones) != T:
            continue
        seq, ph_at = [], []
        for p, a in zip(phones, reps):
            seq += [pdf[p][0] + 1] + [pdf[p][1] + 1] * a
            ph_at += [p] * (a + 1)
        if all(ph_at[t] in proto.allowed_phones[t] for t in range(T)):
            want.add(tuple(seq))
    if not want:
        assert fst is None
        return
    assert accepted(fst, T) == want
    # trimmed: every state lies on an accepted path; numbering is breadth first (times never decrease)
    assert (np.diff(fst["state_times"]) >= 0).all()
    assert set(fst["dst"].tolist()) | {0} == set(range(fst["num_states"]))


def test_supervision_empty_when_too_many_phones():
    tm = kaldi_io.read_transition_model_text(CHAIN_TM)
    proto = cs.alignment_to_proto_supervision(cs.SupervisionOptions(5, 5, 3), [1, 2, 3], [1, 1, 1])
    assert len(proto.allowed_phones) == 1
    assert cs.proto_supervision_to_supervision(mono_tree(), tm, proto) is None


def test_left_biphone_tree_changes_pdfs():
    tree = kaldi_io.read_tree_text(BIPHONE_TREE)
    assert (tree.context_width(), tree.central_position()) == (2, 1)
    assert tree.compute([0, 1], 0) == 0 and tree.compute([2, 1], 1) == 7 and tree.compute([1, 3], 0) == 4
    tm = kaldi_io.read_transition_model_text(CHAIN_TM)
    proto = cs.alignment_to_proto_supervision(cs.SupervisionOptions(0, 0, 1), [1, 1, 2], [2, 2, 1])
    fst = cs.proto_supervision_to_supervision(tree, tm, proto)
    # first phone 1 has no left context (pdfs 0/1), the second one follows phone 1 (still the [0 1] branch),
    # phone 2: pdfs 2/3.  Frames 0..3 allow phone 1, so the boundary between its two instances is free.
    assert accepted(fst, 5) == {(1, 2, 1, 2, 3), (1, 1, 2, 2, 3), (1, 2, 2, 1, 3)}
    proto = cs.alignment_to_proto_supervision(cs.SupervisionOptions(0, 0, 1), [2, 1], [1, 2])
    fst = cs.proto_supervision_to_supervision(tree, tm, proto)
    assert accepted(fst, 3) == {(3, 7, 8)}                    # phone 1 after phone 2: pdfs 6/7


def test_whole_chain_feeds_supervision_object():
    """bin/train_chain.py:262-272 end to end on the host: alignment -> graphs.Supervision (index arrays only)."""
    from pykaldi2_b200 import graphs
    ali_tm = kaldi_io.read_transition_model_text(ALI_TM)
    chain_tm = kaldi_io.read_transition_model_text(CHAIN_TM)
    opts = cs.SupervisionOptions(5, 5, 3)
    ali = [1, 1, 2, 3, 4, 5, 5, 6, 8, 10, 11, 12]
    fst, T = cs.supervision_from_alignment(opts, ali_tm, chain_tm, mono_tree(), ali)
    assert T == 4
    sup = graphs.Supervision(fst, T, 6)
    assert sup.frames_per_sequence == 4 and sup.label_dim == 6 and sup.weight == 1.0


def test_binary_transition_model_and_tree_round_trip(tmp_path):
    """reader.kaldi_io binary forms (Kaldi's io-funcs layout: "\\0B", tokens + space, length-prefixed basic types,
    int32 vectors): what the writers produce the readers parse back to the text-form result; the auto-detecting
    readers take both forms; a GMM-style model file with trailing bytes after </TransitionModel> is accepted."""
    import os
    for name, text in (("ali", ALI_TM), ("chain", CHAIN_TM)):
        tm = kaldi_io.read_transition_model_text(text)
        p = os.path.join(tmp_path, name + ".mdl")
        kaldi_io.write_transition_model_binary(tm, p)
        with open(p, "ab") as f:
            f.write(b"<DIMENSION> \x04\x27\x00\x00\x00")       # whatever follows the transition model is ignored
        assert open(p, "rb").read(2) == b"\0B"
        tb = kaldi_io.read_transition_model(p)
        for k in ("tid2pdf", "tid2phone", "tid2state", "tid_is_self_loop", "tid_is_final"):
            assert np.array_equal(tm[k], tb[k]), (name, k)
        assert tm["tuples"] == tb["tuples"] and tm["pdf_class"] == tb["pdf_class"]
        pt = os.path.join(tmp_path, name + ".txt")
        with open(pt, "w") as f:
            f.write(text)
        assert np.array_equal(kaldi_io.read_transition_model(pt)["tid2pdf"], tm["tid2pdf"])
    # the extended (<Tuples>) format is chosen exactly when forward and self-loop pdfs differ
    raw = open(os.path.join(tmp_path, "chain.mdl"), "rb").read()
    assert b"<Tuples>" in raw and b"<Triples>" in open(os.path.join(tmp_path, "ali.mdl"), "rb").read()
    tree = kaldi_io.read_tree_text(BIPHONE_TREE)
    pt = os.path.join(tmp_path, "tree")
    kaldi_io.write_tree_binary(tree, pt)
    tb = kaldi_io.read_tree(pt)
    assert (tb.context_width(), tb.central_position()) == (2, 1)
    for w in ([0, 1], [1, 1], [2, 1], [3, 1], [1, 2], [0, 3], [2, 3]):
        for c in (0, 1):
            assert tree.compute(w, c) == tb.compute(w, c)
    with pytest.raises(ValueError):
        kaldi_io.read_tree_text(pt)                              # the text-only reader names the conversion command
    # the supervision built from binary assets equals the one from text assets
    opts = cs.SupervisionOptions(5, 5, 3)
    ali = [1, 1, 2, 3, 4, 5, 5, 6, 8, 10, 11, 12]
    a, _ = cs.supervision_from_alignment(opts, kaldi_io.read_transition_model_text(ALI_TM),
                                         kaldi_io.read_transition_model_text(CHAIN_TM), mono_tree(), ali)
    kaldi_io.write_tree_binary(mono_tree(), pt)
    b, _ = cs.supervision_from_alignment(opts, kaldi_io.read_transition_model(os.path.join(tmp_path, "ali.mdl")),
                                         kaldi_io.read_transition_model(os.path.join(tmp_path, "chain.mdl")),
                                         kaldi_io.read_tree(pt), ali)
    assert all(np.array_equal(a[k], b[k]) for k in ("src", "dst", "ilabel", "state_times"))


def test_binary_occupancy_vector(tmp_path):
    import os
    import struct
    p = os.path.join(tmp_path, "final.occs")
    with open(p, "wb") as f:
        f.write(b"\0BFV " + struct.pack("<bi", 4, 3) + struct.pack("<3f", 10.0, 11.5, 3.25))
    assert kaldi_io.read_vector(p).tolist() == [10.0, 11.5, 3.25]
    want = np.log(np.array([10.0, 11.5, 3.25]) / 24.75).astype(np.float32)
    np.testing.assert_allclose(kaldi_io.log_prior_from_occs(p), want, rtol=1e-6)
    np.testing.assert_allclose(kaldi_io.log_prior_from_occs(" [ 10 11.5 3.25 ]"), want, rtol=1e-6)
