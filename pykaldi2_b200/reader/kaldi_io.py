"""Kaldi objects the sequence trainers take from a Kaldi setup (reference bin/train_se.py:164-184 and
bin/train_chain.py:162-181 read them through PyKaldi): the transition model (the head of ``final.mdl`` /
``0.trans_mdl``), the context-dependency tree and the pdf occupancy vector behind the log-prior (``final.occs``), in
text form (``copy-transition-model / copy-tree / copy-vector --binary=false``) or in Kaldi's binary form
(``read_transition_model``, ``read_tree``, ``read_vector`` detect the "\\0B" header).

Only what the hot path needs is extracted from the transition model: transition-id -> pdf and transition-id ->
phone (``TidPdfMap``), plus -- for the chain supervision builder (chain_supervision.py) -- the transition state,
self-loop / final flags of every transition id and the pdf classes of the topology; ``read_tree_text`` reads a
context-dependency tree (``copy-tree --binary=false tree -``) into an evaluable EventMap.  Transition ids are enumerated the way Kaldi's ``TransitionModel::ComputeDerived`` does:
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
    pdf_class = {}                              # phone -> list over hmm states of (forward class, self-loop class) / None
    while i < end_topo:
        if tok[i] == "<TopologyEntry>":
            j = tok.index("<ForPhones>", i) + 1
            phones = []
            while tok[j] != "</ForPhones>":
                phones.append(int(tok[j])); j += 1
            states = []
            classes = []
            k = j + 1
            while tok[k] != "</TopologyEntry>":
                if tok[k] == "<State>":
                    k += 2                      # state index
                    trans = []
                    cls = [None, None]
                    while tok[k] != "</State>":
                        if tok[k] in ("<PdfClass>", "<ForwardPdfClass>", "<SelfLoopPdfClass>"):
                            c = int(tok[k + 1])
                            if tok[k] != "<SelfLoopPdfClass>":
                                cls[0] = c
                            if tok[k] != "<ForwardPdfClass>":
                                cls[1] = c
                            k += 2
                        elif tok[k] == "<Transition>":
                            trans.append((int(tok[k + 1]), float(tok[k + 2]))); k += 3
                        elif tok[k] == "<Final>":
                            k += 2
                        else:
                            raise ValueError("unexpected token %r in a topology state" % tok[k])
                    states.append(trans)
                    classes.append(None if cls[0] is None else (cls[0], cls[1]))
                k += 1
            for p in phones:
                topo[p] = states
                pdf_class[p] = classes
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
    tuples = [(vals[r * width], vals[r * width + 1], vals[r * width + 2],
               vals[r * width + 3] if width == 4 else vals[r * width + 2]) for r in range(n)]
    return _derive(topo, pdf_class, tuples)


def _derive(topo, pdf_class, tuples):
    """TransitionModel::ComputeDerived: per-transition-id tables from the topology and the (phone, hmm-state,
    forward-pdf, self-loop-pdf) tuples."""
    tid2pdf, tid2phone = [-1], [0]              # transition ids start at 1
    tid2state, tid_self, tid_final = [0], [False], [False]
    for r, (phone, hs, fpdf, spdf) in enumerate(tuples):
        for dest, _ in topo[phone][hs]:
            tid2pdf.append(spdf if dest == hs else fpdf)
            tid2phone.append(phone)
            tid2state.append(r + 1)             # transition states are numbered from 1 like Kaldi's
            tid_self.append(dest == hs)
            # TransitionModel::IsFinal: the transition enters the topology's final (non-emitting) state
            tid_final.append(dest < len(topo[phone]) and len(topo[phone][dest]) == 0 or dest >= len(topo[phone]))
    return {"tid2pdf": np.asarray(tid2pdf, np.int32), "tid2phone": np.asarray(tid2phone, np.int32),
            "tid2state": np.asarray(tid2state, np.int32), "tid_is_self_loop": np.asarray(tid_self, bool),
            "tid_is_final": np.asarray(tid_final, bool), "topology": topo, "pdf_class": pdf_class,
            "tuples": list(tuples), "num_pdfs": int(max(tid2pdf)) + 1, "phones": sorted(topo)}


class EventMap(object):
    """Kaldi's decision-tree EventMap (tree/event-map.h) in evaluable form.  An event is a dict key -> value with
    keys 0 .. N-1 = the phones of the context window and key -1 = the pdf class."""

    def __init__(self, kind, key=None, answer=None, table=None, yes_set=None, yes=None, no=None):
        self.kind, self.key, self.answer, self.table = kind, key, answer, table
        self.yes_set, self.yes, self.no = yes_set, yes, no

    def map(self, event):
        node = self
        while True:
            if node is None:
                return None
            if node.kind == "CE":
                return node.answer
            v = event.get(node.key)
            if v is None:
                return None
            if node.kind == "TE":
                node = node.table[v] if 0 <= v < len(node.table) else None
            else:
                node = node.yes if v in node.yes_set else node.no


class ContextDependency(object):
    """ContextDependency(N, P, to_pdf): ``compute(window, pdf_class)`` -> pdf id (tree/context-dep.h Compute)."""

    def __init__(self, context_width, central_position, to_pdf):
        self.N, self.P, self.to_pdf = int(context_width), int(central_position), to_pdf

    def context_width(self):
        return self.N

    def central_position(self):
        return self.P

    def compute(self, window, pdf_class):
        if len(window) != self.N:
            raise ValueError("phone window of %d, tree context width %d" % (len(window), self.N))
        ev = {i: int(p) for i, p in enumerate(window)}
        ev[-1] = int(pdf_class)
        return self.to_pdf.map(ev)

    @classmethod
    def monophone(cls, phone_state_pdf):
        """Tree of a context-independent system from {(phone, pdf_class): pdf}."""
        by_phone = {}
        for (ph, c), pdf in phone_state_pdf.items():
            by_phone.setdefault(ph, {})[c] = pdf
        tab = [None] * (max(by_phone) + 1)
        for ph, d in by_phone.items():
            tab[ph] = EventMap("TE", key=-1, table=[EventMap("CE", answer=d[c]) if c in d else None
                                                     for c in range(max(d) + 1)])
        return cls(1, 0, EventMap("TE", key=0, table=tab))


def read_tree_text(path_or_text):
    """``ContextDependency N P ToPdf <EventMap> EndContextDependency`` in Kaldi's text form
    (CE answer | TE key size ( maps ) | SE key [ yes set ] { yes no } | NULL)."""
    text = path_or_text
    if "ContextDependency" not in path_or_text:
        with open(path_or_text, "rb") as f:
            raw = f.read()
        if raw[:2] == b"\0B":
            raise ValueError("%s is a binary Kaldi tree; convert it with `copy-tree --binary=false tree tree.txt`" % path_or_text)
        text = raw.decode("latin-1")
    tok = text.replace("[", " [ ").replace("]", " ] ").replace("(", " ( ").replace(")", " ) ") \
              .replace("{", " { ").replace("}", " } ").split()
    i = tok.index("ContextDependency")
    N, P = int(tok[i + 1]), int(tok[i + 2])
    if tok[i + 3] != "ToPdf":
        raise ValueError("malformed tree: expected ToPdf")
    pos = [i + 4]

    def parse():
        t = tok[pos[0]]
        pos[0] += 1
        if t == "NULL":
            return None
        if t == "CE":
            a = int(tok[pos[0]]); pos[0] += 1
            return EventMap("CE", answer=a)
        if t == "TE":
            key, size = int(tok[pos[0]]), int(tok[pos[0] + 1])
            if tok[pos[0] + 2] != "(":
                raise ValueError("malformed TE node")
            pos[0] += 3
            table = [parse() for _ in range(size)]
            if tok[pos[0]] != ")":
                raise ValueError("malformed TE node: missing )")
            pos[0] += 1
            return EventMap("TE", key=key, table=table)
        if t == "SE":
            key = int(tok[pos[0]])
            if tok[pos[0] + 1] != "[":
                raise ValueError("malformed SE node")
            pos[0] += 2
            ys = set()
            while tok[pos[0]] != "]":
                ys.add(int(tok[pos[0]])); pos[0] += 1
            if tok[pos[0] + 1] != "{":
                raise ValueError("malformed SE node: missing {")
            pos[0] += 2
            yes = parse()
            no = parse()
            if tok[pos[0]] != "}":
                raise ValueError("malformed SE node: missing }")
            pos[0] += 1
            return EventMap("SE", key=key, yes_set=ys, yes=yes, no=no)
        raise ValueError("unknown EventMap node %r" % t)

    root = parse()
    if tok[pos[0]] != "EndContextDependency":
        raise ValueError("malformed tree: expected EndContextDependency")
    return ContextDependency(N, P, root)


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


def read_vector(path):
    """Kaldi vector file, text (`` [ 1 2 3 ]``) or binary ("\\0B", token FV / DV, int32 size, raw values)."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] != b"\0B":
        return read_vector_text(raw.decode("latin-1"))
    r = _BinReader(raw, 2)
    kind = r.token()
    if kind not in ("FV", "DV"):
        raise ValueError("expected a Kaldi vector (FV / DV), found %r" % kind)
    n = r.int32()
    width, code = (4, "f") if kind == "FV" else (8, "d")
    return np.asarray(struct.unpack_from("<%d%s" % (n, code), raw, r.i), np.float64)


def log_prior_from_occs(path_or_text):
    """log(occs / sum(occs)) (reference bin/train_se.py:183-184); text form, or a binary ``final.occs`` file."""
    import os
    if "[" not in path_or_text and os.path.isfile(path_or_text):
        occ = read_vector(path_or_text)
    else:
        occ = read_vector_text(path_or_text)
    return np.log(occ / occ.sum()).astype(np.float32)


# ------------------------------------------------------------------------------------------- binary forms ----
# Kaldi's binary I/O (base/io-funcs-inl.h): a file starts with "\0B"; tokens are written as text followed by one
# space; a basic type is one length byte (+sizeof for signed, -sizeof for unsigned types) followed by the raw
# little-endian value; an integer vector is one byte sizeof(T), an int32 count and the raw values; a float vector is
# the token "FV", an int32 count (basic type) and raw floats.  The object layouts below follow
# HmmTopology::Write / TransitionModel::Write (hmm/hmm-topology.cc, hmm/transition-model.cc) and
# ContextDependency::Write / EventMap::Write (tree/context-dep.cc, tree/event-map.cc).  Restated from the published
# sources; no Kaldi-written file is available in this sandbox: the readers are pinned to the writers next to them
# (tests/test_chain_supervision.py round trips), not to Kaldi output -- parity unpinned.
import struct


class _BinReader(object):
    def __init__(self, raw, pos=0):
        self.b, self.i = raw, pos

    def token(self):
        while self.b[self.i:self.i + 1].isspace():
            self.i += 1
        j = self.i
        while j < len(self.b) and not self.b[j:j + 1].isspace():
            j += 1
        t = self.b[self.i:j].decode("latin-1")
        self.i = j + 1                                  # the single space after a token
        return t

    def expect(self, tok):
        t = self.token()
        if t != tok:
            raise ValueError("expected %r, found %r at byte %d" % (tok, t, self.i))

    def int32(self):
        n = struct.unpack_from("b", self.b, self.i)[0]
        if abs(n) != 4:
            raise ValueError("expected a 4-byte integer at byte %d (length byte %d)" % (self.i, n))
        v = struct.unpack_from("<i" if n > 0 else "<I", self.b, self.i + 1)[0]
        self.i += 5
        return int(v)

    def float32(self):
        n = struct.unpack_from("b", self.b, self.i)[0]
        if n == 4:
            v = struct.unpack_from("<f", self.b, self.i + 1)[0]
        elif n == 8:
            v = struct.unpack_from("<d", self.b, self.i + 1)[0]
        else:
            raise ValueError("expected a float at byte %d (length byte %d)" % (self.i, n))
        self.i += 1 + n
        return float(v)

    def int_vector(self):
        sz = struct.unpack_from("b", self.b, self.i)[0]
        if sz != 4:
            raise ValueError("expected an int32 vector at byte %d" % self.i)
        n = struct.unpack_from("<i", self.b, self.i + 1)[0]
        v = list(struct.unpack_from("<%di" % n, self.b, self.i + 5)) if n else []
        self.i += 5 + 4 * n
        return v


def _w_token(out, t):
    out.append(t.encode("latin-1") + b" ")


def _w_int32(out, v, unsigned=False):
    out.append(struct.pack("<bI", -4, v) if unsigned else struct.pack("<bi", 4, v))


def _w_float(out, v):
    out.append(struct.pack("<bf", 4, v))


def _w_int_vector(out, v):
    out.append(struct.pack("<bi", 4, len(v)) + struct.pack("<%di" % len(v), *v))


def read_transition_model(path):
    """Text or binary Kaldi transition model (the head of a ``final.mdl`` / ``0.trans_mdl``) -> the dict of
    read_transition_model_text."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] != b"\0B":
        return read_transition_model_text(raw.decode("latin-1"))
    r = _BinReader(raw, 2)
    r.expect("<TransitionModel>")
    r.expect("<Topology>")
    phones = r.int_vector()
    phone2idx = r.int_vector()
    n_entries = r.int32()
    is_hmm = True
    if n_entries == -1:                                 # extended format: self-loop pdf classes follow
        is_hmm = False
        n_entries = r.int32()
    entries, classes = [], []
    for _ in range(n_entries):
        states, cls = [], []
        for _ in range(r.int32()):
            fwd = r.int32()
            slf = fwd if is_hmm else r.int32()
            trans = [(r.int32(), r.float32()) for _ in range(r.int32())]
            states.append(trans)
            cls.append(None if fwd < 0 else (fwd, slf))      # kNoPdf = -1 on non-emitting states
        entries.append(states)
        classes.append(cls)
    r.expect("</Topology>")
    topo = {p: entries[phone2idx[p]] for p in phones}
    pdf_class = {p: classes[phone2idx[p]] for p in phones}
    tag = r.token()
    if tag not in ("<Triples>", "<Tuples>"):
        raise ValueError("expected <Triples> or <Tuples>, found %r" % tag)
    tuples = []
    for _ in range(r.int32()):
        ph, hs, fpdf = r.int32(), r.int32(), r.int32()
        tuples.append((ph, hs, fpdf, r.int32() if tag == "<Tuples>" else fpdf))
    r.expect("</Triples>" if tag == "<Triples>" else "</Tuples>")
    return _derive(topo, pdf_class, tuples)


def write_transition_model_binary(tm, path, log_probs=None):
    """The transition-model head of a Kaldi model file in binary form, from the dict the readers return."""
    phones = sorted(tm["topology"])
    uniq, phone2idx = [], [-1] * (max(phones) + 1)
    for p in phones:
        key = (tm["topology"][p], tm["pdf_class"][p])
        for i, (k, _) in enumerate(uniq):
            if k == key:
                phone2idx[p] = i
                break
        else:
            phone2idx[p] = len(uniq)
            uniq.append((key, p))
    is_hmm = all(c is None or c[0] == c[1] for p in phones for c in tm["pdf_class"][p]) and \
        all(f == s_ for _, _, f, s_ in tm["tuples"])
    out = [b"\0B"]
    _w_token(out, "<TransitionModel>")
    _w_token(out, "<Topology>")
    _w_int_vector(out, phones)
    _w_int_vector(out, phone2idx)
    if not is_hmm:
        _w_int32(out, -1)
    _w_int32(out, len(uniq))
    for (states, classes), _ in uniq:
        _w_int32(out, len(states))
        for trans, c in zip(states, classes):
            _w_int32(out, -1 if c is None else c[0])
            if not is_hmm:
                _w_int32(out, -1 if c is None else c[1])
            _w_int32(out, len(trans))
            for d, pr in trans:
                _w_int32(out, d)
                _w_float(out, pr)
    _w_token(out, "</Topology>")
    _w_token(out, "<Triples>" if is_hmm else "<Tuples>")
    _w_int32(out, len(tm["tuples"]))
    for ph, hs, f, s_ in tm["tuples"]:
        _w_int32(out, ph); _w_int32(out, hs); _w_int32(out, f)
        if not is_hmm:
            _w_int32(out, s_)
    _w_token(out, "</Triples>" if is_hmm else "</Tuples>")
    _w_token(out, "<LogProbs>")
    lp = [0.0] * len(tm["tid2pdf"]) if log_probs is None else list(log_probs)
    _w_token(out, "FV")
    _w_int32(out, len(lp))
    out.append(struct.pack("<%df" % len(lp), *lp))
    _w_token(out, "</LogProbs>")
    _w_token(out, "</TransitionModel>")
    with open(path, "wb") as f:
        f.write(b"".join(out))


def read_tree(path):
    """Text or binary Kaldi context-dependency tree -> ContextDependency."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] != b"\0B":
        return read_tree_text(raw.decode("latin-1"))
    r = _BinReader(raw, 2)
    r.expect("ContextDependency")
    N, P = r.int32(), r.int32()
    r.expect("ToPdf")

    def parse():
        t = r.token()
        if t == "NULL":
            return None
        if t == "CE":
            return EventMap("CE", answer=r.int32())
        if t == "TE":
            key, size = r.int32(), r.int32()
            r.expect("(")
            table = [parse() for _ in range(size)]
            r.expect(")")
            return EventMap("TE", key=key, table=table)
        if t == "SE":
            key = r.int32()
            ys = set(r.int_vector())
            r.expect("{")
            yes, no = parse(), parse()
            r.expect("}")
            return EventMap("SE", key=key, yes_set=ys, yes=yes, no=no)
        raise ValueError("unknown EventMap node %r at byte %d" % (t, r.i))

    root = parse()
    r.expect("EndContextDependency")
    return ContextDependency(N, P, root)


def write_tree_binary(tree, path):
    out = [b"\0B"]
    _w_token(out, "ContextDependency")
    _w_int32(out, tree.N); _w_int32(out, tree.P)
    _w_token(out, "ToPdf")

    def emit(node):
        if node is None:
            _w_token(out, "NULL")
        elif node.kind == "CE":
            _w_token(out, "CE"); _w_int32(out, node.answer)
        elif node.kind == "TE":
            _w_token(out, "TE"); _w_int32(out, node.key); _w_int32(out, len(node.table), unsigned=True)
            _w_token(out, "(")
            for c in node.table:
                emit(c)
            _w_token(out, ")")
        else:
            _w_token(out, "SE"); _w_int32(out, node.key); _w_int_vector(out, sorted(node.yes_set))
            _w_token(out, "{"); emit(node.yes); emit(node.no); _w_token(out, "}")
    emit(tree.to_pdf)
    _w_token(out, "EndContextDependency")
    with open(path, "wb") as f:
        f.write(b"".join(out))
