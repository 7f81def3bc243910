"""FST files -> the dict format of graphs.DenominatorGraph / graphs.Supervision (reader half).

The reference obtains its graphs through PyKaldi (``kaldi_fst.StdVectorFst.read(den_fst_path)``,
bin/train_chain.py:196-202); PyKaldi/OpenFst are not available here, so two on-disk forms are parsed directly:
  * OpenFst's AT&T text form (``fstprint``): ``src dst ilabel olabel [weight]`` per arc, ``state [weight]`` per
    final state, the source of the first line is the start state;
  * OpenFst's binary form of a ``vector`` FST over ``standard`` (tropical, float32) arcs -- what Kaldi writes for
    ``den.fst``: header (magic 2125659606, fst type, arc type, version, flags, properties, start, #states, #arcs,
    optional symbol tables), then per state the final weight, the arc count and the arcs
    (ilabel, olabel, weight, nextstate).
Result: dict(num_states, start, src, dst, ilabel, olabel, weight, final) with arcs sorted by source state and
``final`` = +inf for non-final states (a weight that is absent in the text form is 0).
The binary layout is restated from OpenFst's published file format (fst/fst.h FstHeader, fst/vector-fst.h); no
OpenFst-written file is available in this sandbox, so the reader is pinned by a round trip through
``write_fst_binary`` only (parity with OpenFst itself: unpinned).
"""
import struct

import numpy as np

FST_MAGIC = 2125659606
SYMTAB_MAGIC = 2125658996
_INF = float("inf")


def _finish(num_states, start, arcs, finals):
    arcs.sort(key=lambda a: a[0])
    a = np.asarray(arcs, np.float64).reshape(-1, 5)
    final = np.full(num_states, np.inf, np.float32)
    for s, w in finals.items():
        final[s] = w
    return {"num_states": int(num_states), "start": int(start),
            "src": a[:, 0].astype(np.int32), "dst": a[:, 1].astype(np.int32),
            "ilabel": a[:, 2].astype(np.int32), "olabel": a[:, 3].astype(np.int32),
            "weight": a[:, 4].astype(np.float32), "final": final}


def read_fst_text(path_or_lines):
    """AT&T text form (what ``fstprint`` writes, integer labels)."""
    if isinstance(path_or_lines, str):
        with open(path_or_lines) as f:
            lines = f.read().splitlines()
    else:
        lines = list(path_or_lines)
    arcs, finals, start, nstates = [], {}, None, 0
    for line in lines:
        p = line.split()
        if not p:
            continue
        if len(p) >= 4:
            s, d, il, ol = int(p[0]), int(p[1]), int(p[2]), int(p[3])
            w = float(p[4]) if len(p) > 4 else 0.0
            arcs.append((s, d, il, ol, w))
            nstates = max(nstates, s + 1, d + 1)
        elif len(p) <= 2:
            s = int(p[0])
            w = float(p[1]) if len(p) > 1 else 0.0
            if w != _INF:
                finals[s] = w
            nstates = max(nstates, s + 1)
        else:
            raise ValueError("bad FST text line: %r" % line)
        if start is None:
            start = s
    if start is None:
        raise ValueError("empty FST")
    return _finish(nstates, start, arcs, finals)


def write_fst_text(fst, path=None):
    lines = []
    order = np.argsort(np.asarray(fst["src"]) != fst["start"], kind="stable")     # the start state's arcs first
    ol = fst.get("olabel", fst["ilabel"])
    for k in order:
        lines.append("%d %d %d %d %.9g" % (fst["src"][k], fst["dst"][k], fst["ilabel"][k], ol[k], fst["weight"][k]))
    if not len(order) or fst["src"][order[0]] != fst["start"]:
        lines.insert(0, "%d" % fst["start"] if np.isfinite(fst["final"][fst["start"]]) else "%d inf" % fst["start"])
    for s, w in enumerate(np.asarray(fst["final"])):
        if np.isfinite(w):
            lines.append("%d %.9g" % (s, w))
    text = "\n".join(lines) + "\n"
    if path:
        with open(path, "w") as f:
            f.write(text)
    return text


class _Cursor(object):
    def __init__(self, blob):
        self.b, self.i = blob, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.i)
        self.i += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def string(self):
        n = self.take("i")
        s = self.b[self.i:self.i + n].decode()
        self.i += n
        return s


def _skip_symbol_table(c):
    if c.take("i") != SYMTAB_MAGIC:
        raise ValueError("bad symbol table magic")
    c.string()                      # name
    c.take("q")                     # available key
    n = c.take("q")
    for _ in range(n):
        c.string()
        c.take("q")


def read_fst_binary(path_or_bytes):
    blob = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    c = _Cursor(blob)
    if c.take("i") != FST_MAGIC:
        raise ValueError("not an OpenFst binary file (bad magic)")
    fst_type, arc_type = c.string(), c.string()
    version, flags = c.take("i"), c.take("i")
    c.take("Q")                     # properties
    start, nstates, _narcs = c.take("q"), c.take("q"), c.take("q")
    if fst_type != "vector" or arc_type != "standard":
        raise ValueError("only 'vector' FSTs over 'standard' arcs are read (got %s / %s); convert with fstconvert" % (fst_type, arc_type))
    if flags & 4:
        raise ValueError("aligned OpenFst files are not supported")
    if flags & 1:
        _skip_symbol_table(c)
    if flags & 2:
        _skip_symbol_table(c)
    arcs, finals = [], {}
    for s in range(nstates):
        w = c.take("f")
        if w != _INF:
            finals[s] = w
        for _ in range(c.take("q")):
            il, ol, aw, d = c.take("iifi")
            arcs.append((s, d, il, ol, aw))
    del version
    return _finish(nstates, start, arcs, finals)


def write_fst_binary(fst, path=None):
    src = np.asarray(fst["src"])
    ol = fst.get("olabel", fst["ilabel"])
    S = int(fst["num_states"])

    def string(s):
        return struct.pack("<i", len(s)) + s.encode()
    out = [struct.pack("<i", FST_MAGIC), string("vector"), string("standard"), struct.pack("<iiQ", 2, 0, 0),
           struct.pack("<qqq", int(fst["start"]), S, len(src))]
    off = np.searchsorted(src, np.arange(S + 1))
    for s in range(S):
        out.append(struct.pack("<fq", float(fst["final"][s]), int(off[s + 1] - off[s])))
        for k in range(off[s], off[s + 1]):
            out.append(struct.pack("<iifi", int(fst["ilabel"][k]), int(ol[k]), float(fst["weight"][k]), int(fst["dst"][k])))
    blob = b"".join(out)
    if path:
        with open(path, "wb") as f:
            f.write(blob)
    return blob


def read_fst(path):
    """Binary if the file starts with OpenFst's magic number, else AT&T text."""
    with open(path, "rb") as f:
        head = f.read(4)
    if len(head) == 4 and struct.unpack("<i", head)[0] == FST_MAGIC:
        return read_fst_binary(path)
    return read_fst_text(path)
