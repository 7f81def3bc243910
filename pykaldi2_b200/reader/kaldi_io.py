"""Text-form Kaldi objects the sequence trainers take from a Kaldi setup (reference bin/train_se.py:164-184 reads
them through PyKaldi): the transition model (``copy-transition-model --binary=false final.mdl -``) and the pdf
occupancy vector behind the log-prior (``final.occs`` in text form).

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
    tid2pdf, tid2phone = [-1], [0]              # transition ids start at 1
    tid2state, tid_self, tid_final = [0], [False], [False]
    for r in range(n):
        phone, hs, fpdf = vals[r * width], vals[r * width + 1], vals[r * width + 2]
        spdf = vals[r * width + 3] if width == 4 else fpdf
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
            "num_pdfs": int(max(tid2pdf)) + 1, "phones": sorted(topo)}


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


def log_prior_from_occs(path_or_text):
    """log(occs / sum(occs)) (reference bin/train_se.py:183-184)."""
    occ = read_vector_text(path_or_text)
    return np.log(occ / occ.sum()).astype(np.float32)
