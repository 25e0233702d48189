"""Protocol-2 pickle of a TSSB in the form the reference's readers expect (util2.py:250-262, TreeReader util2.py:264-343,
pwgsresults/result_generator.py:11-12), produced several times faster than `pickle.dumps`.

Readable by Python 2 / numpy 1 (`compat_pickler`): numpy >= 2 pickles its arrays and scalars through
`numpy._core.multiarray._reconstruct / scalar`, a module numpy 1.x does not have.  The stream therefore opens with a hand-written
preamble that defines those two callables under their numpy-1 names (`numpy.core.multiarray`, still importable under numpy 2)
at memo slots 0 and 1, and the pickler starts with the same two entries in its memo, so every array and scalar in the stream is a
`BINGET 0/1` instead of a GLOBAL.  Everything else is protocol-2 material both sides know (tests/test_pickle_compat.py: GLOBAL
whitelist, opcode whitelist, and the reference's own TreeReader run on the archive).

What costs in the plain pickler is the list of D `Datum` objects: ~15 attributes each, all constant over the chain except the
reference to the datum's node.  So the tree is pickled in two parts:

  * everything but `tssb.data` goes through the ordinary C pickler (nodes, sticks, the assignment list of node references, ...)
    with a placeholder string where the data list belongs; the pickler's memo then tells under which index every node object
    and the TSSB itself can be referenced;
  * the data list is written by hand from cached byte fragments, one per datum (its constant attributes, pickled once with a
    memo-less pickler), joined with a `LONG_BINGET <node>` per datum.  Memo indices of the hand-written part start at 2^16, above
    the ordinary pickler's (a tree with more memo entries than that falls back).  Datums that other datums point to (the CNVs of an SSM's `cnv` list) are defined in a preamble
    (`... POP`) so that every reference follows its definition.

The join of the ~3 D byte strings of the second part is done by host code of the library (pwgs_pickle_join,
csrc/pwgs_host_pickle.cpp) from the assignment array the sweeps maintain (backend.assignment_array) straight into the result,
which is then a `bytearray` (what zipfile.writestr and pickle.loads take as they take bytes); the Python join remains for
trees the shortcut does not cover and is byte-identical (tests/test_fastpickle.py).

The result unpickles (plain `pickle.load`) to the same object graph: same attributes, `datum.node` the very node
object of `tssb.assignments`, `datum.tssb` the tree, `cnv` links pointing at the CNV datums of the same list.  Anything
unexpected (an attribute that is not a plain value, a reference cycle between datums) falls back to the ordinary pickler."""
import io
import pickle
import struct
import threading

_PLACEHOLDER = "\x00pwgs-b200 data list goes here\x00"
_BASE = 1 << 16                                          # first memo index of the hand-written part: above the ordinary pickler's
                                                         # (checked), small because Python 3's unpickler keeps its memo in an array

_MARK, _POP, _STOP = b"(", b"0", b"."
_EMPTY_LIST, _APPENDS, _EMPTY_DICT, _SETITEMS, _BUILD = b"]", b"e", b"}", b"u", b"b"
_EMPTY_TUPLE, _NEWOBJ, _NONE, _TUPLE3 = b")", b"\x81", b"N", b"\x87"


def _put(i):
    return b"r" + struct.pack("<I", i)                  # LONG_BINPUT


def _get(i):
    return (b"h" + struct.pack("<B", i)) if i < 256 else (b"j" + struct.pack("<I", i))     # BINGET / LONG_BINGET


def _numpy_reducers():
    try:
        from numpy._core import multiarray as m          # numpy >= 2
    except ImportError:                                  # pragma: no cover - numpy 1.x names them as the readers do
        from numpy.core import multiarray as m
    return m._reconstruct, m.scalar


# GLOBAL numpy.core.multiarray._reconstruct, BINPUT 0, POP; GLOBAL numpy.core.multiarray.scalar, BINPUT 1, POP
_NP_PREAMBLE = b"cnumpy.core.multiarray\n_reconstruct\nq\x000cnumpy.core.multiarray\nscalar\nq\x010"


def compat_pickler(buf):
    """A protocol-2 pickler whose output, passed through `compat_finish`, names numpy's reconstructors the numpy-1 way."""
    pk = pickle.Pickler(buf, protocol=2)
    pk.memo = {id(f): (i, f) for i, f in enumerate(_numpy_reducers())}
    return pk


def compat_finish(stream):
    assert stream[:2] == b"\x80\x02"
    return b"".join([stream[:2], _NP_PREAMBLE, stream[2:]])


def compat_dumps(obj):
    """pickle.dumps(obj, protocol=2), readable by Python 2 / numpy 1."""
    buf = io.BytesIO()
    compat_pickler(buf).dump(obj)
    return compat_finish(buf.getvalue())


_tls = threading.local()                                 # one memo-less pickler per thread: chains may share a process
_small = {}


def _plain(value):
    """Pickle body of a plain value (numbers, strings, lists / tuples of them) without memo traffic, header or STOP."""
    if type(value) is int and -1 <= value < 256:         # the common case, cached (idempotent: safe to race on)
        body = _small.get(value)
        if body is not None:
            return body
    buf = getattr(_tls, "buf", None)
    if buf is None:
        buf = _tls.buf = io.BytesIO()
        _tls.pickler = pickle.Pickler(buf, protocol=2)
        _tls.pickler.fast = True                         # no memo: no PUT opcodes that could collide with the main stream
    buf.seek(0)
    buf.truncate()
    _tls.pickler.dump(value)
    body = buf.getvalue()
    assert body[:2] == b"\x80\x02" and body[-1:] == _STOP
    body = body[2:-1]
    if type(value) is int and -1 <= value < 256:
        _small[value] = body
    return body


def _is_plain(v, depth=0):
    if v is None or isinstance(v, (bool, int, float, str)):
        return True
    if isinstance(v, (list, tuple)) and depth < 3:
        return all(_is_plain(x, depth + 1) for x in v)
    return False


class TreePickler(object):
    """Caches the constant fragment of every datum of one chain's data list."""

    VARIABLE = ("node", "tssb")

    def __init__(self, data):
        self.data = data
        self.ok = True
        self.keys = {}                                   # attribute name -> memo index
        pos = {id(d): i for i, d in enumerate(data)}
        self.pos = pos
        try:
            self.cls_global = b"c" + type(data[0]).__module__.encode() + b"\n" + type(data[0]).__name__.encode() + b"\n"
            referenced = set()
            frags = [self._fragment(i, referenced) for i in range(len(data))]
            self._key("tssb")
            self.frags = frags
            self.referenced = sorted(referenced)
            self._native_cache()
        except (ValueError, KeyError, TypeError, AttributeError):
            self.ok = False

    def _fragment(self, i, referenced):
        """Everything of datum i up to the value of its `node` attribute."""
        data, d = self.data, self.data[i]
        if type(d) is not type(data[0]):
            raise ValueError("mixed datum classes")
        parts = [_get(_BASE), _EMPTY_TUPLE, _NEWOBJ, _put(_BASE + 1 + i), _EMPTY_DICT, _MARK]
        for key, v in d.__dict__.items():
            if key in self.VARIABLE:
                continue
            parts.append(self._key(key))
            if key == "cnv":
                parts.append(_EMPTY_LIST)
                if v:
                    parts.append(_MARK)
                    for other, c1, c2 in v:
                        j = self.pos[id(other)]
                        if data[j].cnv:
                            raise ValueError("linked datum has links of its own")
                        referenced.add(j)
                        parts += [_get(_BASE + 1 + j), _plain(c1), _plain(c2), _TUPLE3]
                    parts.append(_APPENDS)
            elif _is_plain(v):
                parts.append(_plain(v))
            else:
                raise ValueError("attribute %s of a datum is not a plain value" % key)
        parts.append(self._key("node"))
        return b"".join(parts)

    def _native_cache(self):
        """The list part as one byte string + offsets for pwgs_pickle_join (csrc/pwgs_host_pickle.cpp): the fragment of every
        datum, a back-reference for those the preamble has already defined."""
        import numpy as np
        parts = list(self.frags)
        for j in self.referenced:
            parts[j] = _get(_BASE + 1 + j)
        self.blob = b"".join(parts)
        self.off = np.zeros(len(parts) + 1, dtype=np.int64)
        np.cumsum([len(p) for p in parts], out=self.off[1:])
        self.standalone = np.array(self.referenced, dtype=np.int64)

    def _join_native(self, tssb, memo, tail, prefix, suffix):
        """prefix + b"".join of [fragment, node reference, tail] over the data + suffix as one bytearray, the join done by the
        library's host code straight into the result; None when the shortcut does not apply (the caller then joins in Python)."""
        import ctypes as C
        import numpy as np
        from .. import _lib
        from .backend import assignment_array
        try:
            lib = _lib.load()
        except _lib.PwgsError:
            return None
        data, assignments = self.data, tssb.assignments
        n = len(data)
        if len(assignments) != n or any(data[i].node is not assignments[i] for i in {0, n // 2, n - 1}):
            return None                                  # datum.node is what the format stores; it must be the assignment
        nodes = tssb.get_nodes()
        if any(id(nd) not in memo for nd in nodes):
            return None
        nod = assignment_array(tssb, nodes)
        if nod is None:
            return None
        refs = np.zeros((len(nodes), 8), dtype=np.uint8)
        for k, nd in enumerate(nodes):
            r = _get(memo[id(nd)][0])
            refs[k, 0] = len(r)
            refs[k, 1:1 + len(r)] = np.frombuffer(r, dtype=np.uint8)
        ref_of = np.ascontiguousarray(nod, dtype=np.int32)
        if len(self.standalone):
            ref_of = ref_of.copy()
            ref_of[self.standalone] = -1
        p0 = len(prefix)
        cap = len(self.blob) + n * (5 + len(tail))
        out = bytearray(p0 + cap + len(suffix))
        out[:p0] = prefix
        window = (C.c_char * len(out)).from_buffer(out)
        w = lib.pwgs_pickle_join(n, self.blob, self.off.ctypes.data, ref_of.ctypes.data_as(C.POINTER(C.c_int32)), len(nodes),
                                 refs.ctypes.data, tail, len(tail), C.byref(window, p0), cap)
        del window                                       # the export must be gone before the bytearray is cut to size
        if w < 0:
            return None
        out[p0 + w:p0 + w + len(suffix)] = suffix
        del out[p0 + w + len(suffix):]
        return out

    def _fresh(self):
        """The cache assumes the datums' own attributes never change; spot-check a few."""
        n = len(self.data)
        try:
            n_keys = len(self.keys)
            ok = all(self._fragment(i, set()) == self.frags[i] for i in {0, n // 2, n - 1})
            return ok and len(self.keys) == n_keys
        except (ValueError, KeyError, TypeError, AttributeError):
            return False

    def _key(self, name):
        if name not in self.keys:
            self.keys[name] = _BASE + 1 + len(self.data) + len(self.keys)
        return _get(self.keys[name])

    def dumps(self, tssb, root=None):
        """Pickle of `root` (default: the tree itself), an object graph that contains the tree -- e.g. the chain's state dict."""
        root = tssb if root is None else root
        data = tssb.data
        if not self.ok or data is not self.data or not self._fresh():
            return compat_dumps(root)
        # 1. everything but the data list -- which goes last, so that every node is defined before the list refers to it
        buf = io.BytesIO()
        pk = compat_pickler(buf)
        saved = tssb.__dict__
        stub = dict(saved)
        stub.pop("data")
        stub["data"] = _PLACEHOLDER
        tssb.__dict__ = stub
        try:
            pk.dump(root)
        finally:
            tssb.__dict__ = saved
        main = compat_finish(buf.getbuffer())
        memo = pk.memo.copy()                            # id(obj) -> (index, obj)
        marker = _plain(_PLACEHOLDER)
        at = main.find(marker)
        if at < 0 or main.find(marker, at + 1) >= 0 or id(tssb) not in memo or len(memo) >= _BASE:
            return compat_dumps(root)
        end = at + len(marker)
        if main[end:end + 1] == b"q":                    # the placeholder string's own BINPUT / LONG_BINPUT
            end += 2
        elif main[end:end + 1] == b"r":
            end += 5
        # 2. the data list: per datum [fragment, reference to its node, tail], assembled with slice assignments (no Python-level
        #    work per datum beyond reading its `node`)
        tail = self._key("tssb") + _get(memo[id(tssb)][0]) + _SETITEMS + _BUILD
        out = [self.cls_global, _put(_BASE), _POP]
        for name, idx in self.keys.items():
            out += [_plain(name), _put(idx), _POP]
        frags = self.frags
        view = memoryview(main)
        try:
            for j in self.referenced:                    # define the linked-to datums first
                nd = data[j].node
                out += [frags[j], _NONE if nd is None else _get(memo[id(nd)][0]), tail, _POP]
        except (KeyError, AttributeError):               # a datum's node is not part of the tree
            return compat_dumps(root)
        out += [_EMPTY_LIST, _MARK]
        whole = self._join_native(tssb, memo, tail, b"".join([view[:at]] + out), b"".join([_APPENDS, view[end:]]))
        if whole is not None:
            return whole
        noderef = {id(None): _NONE}
        for oid, (idx, obj) in memo.items():
            noderef[oid] = _get(idx)
        try:
            refs = [noderef[id(d.node)] for d in data]
        except (KeyError, AttributeError):
            return compat_dumps(root)
        n = len(data)
        body = [tail] * (3 * n)
        body[0::3] = frags
        body[1::3] = refs
        for j in self.referenced:                        # already defined: a reference instead of a second definition
            body[3 * j], body[3 * j + 1], body[3 * j + 2] = _get(_BASE + 1 + j), b"", b""
        out += body
        out.append(_APPENDS)
        return b"".join([view[:at]] + out + [view[end:]])       # one allocation, one copy of every part


def dumps(tssb, root=None):
    """pickle.dumps(root or tssb, protocol=2) for a chain's tree (or an object holding it, like the chain's state dict), with the
    per-chain cache kept on the tree."""
    tp = tssb.__dict__.get("_tree_pickler")
    if tp is None or tp.data is not tssb.data:
        tp = TreePickler(tssb.data)
        tssb.__dict__["_tree_pickler"] = tp
    return tp.dumps(tssb, root)
