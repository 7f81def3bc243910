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
