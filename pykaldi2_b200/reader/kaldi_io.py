"""Text-form Kaldi objects the sequence trainers take from a Kaldi setup (reference bin/train_se.py:164-184 reads
them through PyKaldi): the transition model (``copy-transition-model --binary=false final.mdl -``) and the pdf
occupancy vector behind the log-prior (``final.occs`` in text form).

Only what the hot path needs is extracted from the transition model: transition-id -> pdf and transition-id ->
phone (``TidPdfMap``).  Transition ids are enumerated the way Kaldi's ``TransitionModel::ComputeDerived`` does:
tuples (phone, hmm-state, forward-pdf[, self-loop-pdf]) in file order, each owning as many consecutive ids
(starting at 1) as its hmm-state has transitions in the phone's topology entry; an id whose transition returns
to its own hmm-state is a self-loop and maps to the self-loop pdf.  Binary models are rejected with the command
that converts them.  Restated from Kaldi's documented text format -- no Kaldi file is available in this sandbox:
parity unpinned, covered by a hand-built model in tests/test_host.py.
"""
import re

import numpy as np


def _tokens(text):
    return text.replace("[", " [ ").replace("]", " ] ").split()


def read_transition_model_text(path_or_text):
    text = path_or_text
    if "\n" not in path_or_text and "<" not in path_or_text:
        with open(path_or_text, "rb") as f:
            raw = f.read()
        if raw[:2] == b"\0B":
            raise ValueError("%s is a binary Kaldi model; convert it with "
                             "`copy-transition-model --binary=false final.mdl final.mdl.txt`" % path_or_text)
        text = raw.decode("latin-1")
    tok = _tokens(text)
    if "<TransitionModel>" not in tok:
        raise ValueError("no <TransitionModel> in the input")
    i = tok.index("<Topology>")
    end_topo = tok.index("</Topology>")
    topo = {}                                   # phone -> list over hmm states of [(dest, prob), ...]
    while i < end_topo:
        if tok[i] == "<TopologyEntry>":
            j = tok.index("<ForPhones>", i) + 1
            phones = []
            while tok[j] != "</ForPhones>":
                phones.append(int(tok[j])); j += 1
            states = []
            k = j + 1
            while tok[k] != "</TopologyEntry>":
                if tok[k] == "<State>":
                    k += 2                      # state index
                    trans = []
                    while tok[k] != "</State>":
                        if tok[k] in ("<PdfClass>", "<ForwardPdfClass>", "<SelfLoopPdfClass>"):
                            k += 2
                        elif tok[k] == "<Transition>":
                            trans.append((int(tok[k + 1]), float(tok[k + 2]))); k += 3
                        elif tok[k] == "<Final>":
                            k += 2
                        else:
                            raise ValueError("unexpected token %r in a topology state" % tok[k])
                    states.append(trans)
                k += 1
            for p in phones:
                topo[p] = states
            i = k
        i += 1
    if "<Triples>" in tok:
        t0, width, endtag = tok.index("<Triples>"), 3, "</Triples>"
    elif "<Tuples>" in tok:
        t0, width, endtag = tok.index("<Tuples>"), 4, "</Tuples>"
    else:
        raise ValueError("no <Triples>/<Tuples> in the transition model")
    n = int(tok[t0 + 1])
    vals = [int(v) for v in tok[t0 + 2:t0 + 2 + n * width]]
    if tok[t0 + 2 + n * width] != endtag:
        raise ValueError("malformed %s section" % endtag)
    tid2pdf, tid2phone = [-1], [0]              # transition ids start at 1
    for r in range(n):
        phone, hs, fpdf = vals[r * width], vals[r * width + 1], vals[r * width + 2]
        spdf = vals[r * width + 3] if width == 4 else fpdf
        for dest, _ in topo[phone][hs]:
            tid2pdf.append(spdf if dest == hs else fpdf)
            tid2phone.append(phone)
    return {"tid2pdf": np.asarray(tid2pdf, np.int32), "tid2phone": np.asarray(tid2phone, np.int32),
            "num_pdfs": int(max(tid2pdf)) + 1, "phones": sorted(topo)}


def read_vector_text(path_or_text):
    """Kaldi text vector `` [ 1 2 3 ]`` (e.g. final.occs) -> float64 array."""
    text = path_or_text
    if "[" not in path_or_text:
        with open(path_or_text, "rb") as f:
            raw = f.read()
        if raw[:2] == b"\0B":
            raise ValueError("%s is a binary Kaldi vector; convert it with `copy-vector --binary=false`" % path_or_text)
        text = raw.decode("latin-1")
    m = re.search(r"\[([^\]]*)\]", text)
    if not m:
        raise ValueError("no [ ... ] vector found")
    return np.asarray(m.group(1).split(), dtype=np.float64)


def log_prior_from_occs(path_or_text):
    """log(occs / sum(occs)) (reference bin/train_se.py:183-184)."""
    occ = read_vector_text(path_or_text)
    return np.log(occ / occ.sum()).astype(np.float32)
