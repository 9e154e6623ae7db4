"""`.awkd` files: the reference's zip-of-blobs container (awkward0/persist.py:282-409 serialize, 497-588
deserialize, 661-717 save / load) for the arrays of the JaggedArray hot path -- the ingest side of the path
(SURVEY.md section 8f-3): a JaggedArray is stored as (counts, content) blobs, integers of 8 KiB and more
zlib-compressed (persist.py:26-28), and rebuilt by ``JaggedArray.fromcounts`` (jagged.py:289-302).

Loading: the blobs are inflated on the host (zlib has no place on the device), copied host -> HBM once, and the
offsets scan (``jg_counts2offsets``) is the first device step -- ``fromcounts`` exactly as the schema says.
Saving: the inverse, counts from ``jg_offsets2counts``.  Files written here load in the reference
(``awkward0.load``) and files written by the reference load here; tests/golden/*.awkd are written by the
UNMODIFIED reference serializer (oracle/make_golden_awkd.py).

The schema is JSON with a whitelist of constructors; only the ones the hot path can represent are accepted
(JaggedArray.fromcounts, JaggedArray(starts, stops, content), Table.frompairs / fromview, numpy.frombuffer,
zlib.decompress);
anything else (ChunkedArray, VirtualArray, ObjectArray, pickled python objects ...) raises NotImplementedError --
nothing is ever evaluated on the strength of a file's say-so and there is no CPU fallback.

The host half (``read_tree`` / ``write_tree``) works on plain numpy arrays and runs without a GPU.
"""
import json
import os
import zipfile
import zlib
from collections import OrderedDict

import numpy

FORMAT_VERSION = "0.15.5"          # the reference release whose schema this reads and writes
COMPRESS_MINSIZE = 8192            # persist.py:26-28: bool / integer blobs of at least this many bytes are deflated


# ---- host trees ---------------------------------------------------------------------------------------------------
class HostJagged(object):
    """fromcounts form when `counts` is given, general form (starts, stops, content) otherwise."""

    def __init__(self, content, counts=None, starts=None, stops=None):
        self.content, self.counts, self.starts, self.stops = content, counts, starts, stops


class HostTable(object):
    def __init__(self, columns, rowstart=0):
        self.columns, self.rowstart = OrderedDict(columns), rowstart


def _is_call(node, *spec):
    return isinstance(node, dict) and list(node.get("call", ())) == list(spec)


def _json2dtype(obj):
    if not isinstance(obj, str):
        raise NotImplementedError("structured / sub-array dtypes in .awkd files are outside the GPU hot path")
    return numpy.dtype(obj)


def _view_rows(column, start, step, length):
    if isinstance(column, numpy.ndarray):
        return column[start:start + step * length:step] if length > 0 else column[:0]
    if isinstance(column, HostTable):
        return HostTable([(n, _view_rows(c, start, step, length)) for n, c in column.columns.items()], 0)
    raise NotImplementedError("views of nested columns in .awkd files are outside the GPU hot path")


def read_tree(storage, name="", schemasuffix=".json"):
    """storage: mapping name -> bytes (a zip member reader).  Returns a tree of numpy arrays / HostJagged / HostTable."""
    schema = json.loads(bytes(storage[name + schemasuffix]).decode("ascii"))
    if "awkward0" not in schema and "awkward" not in schema:
        raise ValueError("JSON object is not an Awkward Array schema (missing 'awkward' or 'awkward0' field). schema is: {}".format(schema))
    prefix = schema.get("prefix", "")
    seen = {}

    def blob(node):
        if "read" in node:
            return storage[node["read"] if node.get("absolute", False) else prefix + node["read"]]
        if _is_call(node, "zlib", "decompress"):
            return zlib.decompress(blob(node["args"][0]))
        raise NotImplementedError("unsupported buffer source in .awkd schema: {0}".format(node.get("call", list(node))))

    def plain(node):
        if "ref" in node:                   # the reference interns by object identity: equal small integers are refs
            return seen[node["ref"]]
        if "json" in node:
            if "id" in node:
                seen[node["id"]] = node["json"]
            return node["json"]
        raise NotImplementedError("unsupported value in .awkd schema: {0}".format(list(node)))

    def unfill(node):
        if not isinstance(node, dict):
            raise ValueError("unrecognized JSON object: {0}".format(repr(node)))
        if "ref" in node:
            if node["ref"] not in seen:
                raise NotImplementedError("forward references (cyclic arrays) are outside the GPU hot path")
            return seen[node["ref"]]
        if _is_call(node, "awkward0", "numpy", "frombuffer") or _is_call(node, "awkward", "numpy", "frombuffer"):
            buf, dt, count = node["args"]
            out = numpy.frombuffer(blob(buf), dtype=_json2dtype(dt["dtype"]), count=plain(count))
        elif _is_call(node, "awkward0", "JaggedArray", "fromcounts") or _is_call(node, "awkward", "JaggedArray", "fromcounts"):
            counts, content = [unfill(x) for x in node["args"]]
            out = HostJagged(content, counts=counts)
        elif _is_call(node, "awkward0", "JaggedArray") or _is_call(node, "awkward", "JaggedArray"):
            starts, stops, content = [unfill(x) for x in node["args"]]
            out = HostJagged(content, starts=starts, stops=stops)
        elif _is_call(node, "awkward0", "Table", "frompairs") or _is_call(node, "awkward", "Table", "frompairs"):
            pairs, rowstart = node["args"]
            out = HostTable([(n, unfill(x)) for n, x in pairs["pairs"]], plain(rowstart))
        elif _is_call(node, "awkward0", "Table", "fromview") or _is_call(node, "awkward", "Table", "fromview"):
            # table.py:300-310: rows start, start + step, ... (length of them) of the base table -- what slicing a
            # Table leaves behind (JaggedArray.__awkward_serialize__ slices its content, jagged.py:294)
            view, base = node["args"]
            if "tuple" not in view:
                raise NotImplementedError("Table views through an index array are outside the GPU hot path")
            start, step, length = [plain(x) for x in view["tuple"]]
            base = unfill(base)
            out = HostTable([(n, _view_rows(c, start, step, length)) for n, c in base.columns.items()], 0)
        elif "call" in node:
            raise NotImplementedError("{0} in an .awkd file is outside the GPU hot path (SURVEY.md section 8a); no CPU fallback".format(".".join(node["call"])))
        else:
            raise NotImplementedError("unsupported node in .awkd schema: {0}".format(list(node)))
        if "id" in node:
            seen[node["id"]] = out
        return out

    return unfill(schema["schema"])


def write_tree(storage, tree, name="", delimiter="-", suffix=".raw", schemasuffix=".json", compression=True):
    """The reference's BlobSerializer (persist.py:411-495) for the node kinds above; ids are handed out in the
    order the reference hands them out, so the files are the same byte for byte up to zlib's level."""
    prefix = name + delimiter if name else ""
    counter = [0]

    def next_id():
        counter[0] += 1
        return counter[0] - 1

    def array(a, ident):
        a = numpy.ascontiguousarray(a)
        if a.ndim != 1:
            raise NotImplementedError("only one-dimensional arrays are written")
        raw = a.tobytes()
        key = str(ident) + suffix
        deflate = compression and a.nbytes >= COMPRESS_MINSIZE and a.dtype.kind in "biu"
        storage[prefix + key] = zlib.compress(raw) if deflate else raw
        buf = {"call": ["zlib", "decompress"], "args": [{"read": key}]} if deflate else {"read": key}
        return {"call": ["awkward0", "numpy", "frombuffer"], "args": [buf, {"dtype": str(a.dtype)}, {"json": len(a), "id": next_id()}], "id": ident}

    def fill(x):
        ident = next_id()
        if isinstance(x, numpy.ndarray):
            return array(x, ident)
        if isinstance(x, HostJagged):
            if x.counts is not None:
                return {"call": ["awkward0", "JaggedArray", "fromcounts"], "args": [fill(x.counts), fill(x.content)], "id": ident}
            return {"call": ["awkward0", "JaggedArray"], "args": [fill(x.starts), fill(x.stops), fill(x.content)], "id": ident}
        if isinstance(x, HostTable):
            pairs = [[n, fill(c)] for n, c in x.columns.items()]
            return {"call": ["awkward0", "Table", "frompairs"], "args": [{"pairs": pairs}, {"json": x.rowstart}], "id": ident}
        raise TypeError("cannot write {0} to an .awkd file".format(type(x)))

    schema = {"awkward0": FORMAT_VERSION, "schema": fill(tree)}
    if prefix:
        schema["prefix"] = prefix
    storage[name + schemasuffix] = json.dumps(schema).encode("ascii")
    return schema


# ---- zip container (persist.py:661-717, 720-767) ----------------------------------------------------------------
class _ZipWriter(object):
    def __init__(self, f):
        self.f = f

    def __setitem__(self, where, what):
        self.f.writestr(where, what, compress_type=zipfile.ZIP_STORED)


class _ZipReader(object):
    def __init__(self, f):
        self.f = f

    def __getitem__(self, where):
        return self.f.read(where)


def _path(file):
    if hasattr(file, "__fspath__"):
        file = os.fspath(file)
    return file


def save_host(file, trees, mode="a", **options):
    """trees: dict name -> host tree.  Same member layout as awkward0.save."""
    file = _path(file)
    if isinstance(file, str) and not file.endswith(".awkd"):
        file = file + ".awkd"
    names = list(trees)
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            if names[i].startswith(names[j]) or names[j].startswith(names[i]):
                raise KeyError("cannot write both {0} and {1} to zipfile because one is a prefix of the other".format(repr(names[i]), repr(names[j])))
    with zipfile.ZipFile(file, mode=mode, compression=zipfile.ZIP_STORED) as f:
        existing = f.namelist()
        for name in names:
            if any(n.startswith(name) for n in existing):
                raise KeyError("cannot add {0} to zipfile because the following already exist: {1}".format(repr(name), ", ".join(repr(n) for n in existing if n.startswith(name))))
        for name, tree in trees.items():
            write_tree(_ZipWriter(f), tree, name=name, **options)


def load_host(file, name=None):
    """name=None: the single unnamed array if that is all the file holds, else an OrderedDict name -> tree."""
    with zipfile.ZipFile(_path(file), mode="r") as f:
        members = [n[:-5] for n in f.namelist() if n.endswith(".json")]
        reader = _ZipReader(f)
        if name is not None:
            return read_tree(reader, name)
        if members == [""]:
            return read_tree(reader, "")
        return OrderedDict((n, read_tree(reader, n)) for n in members)


# ---- device side -------------------------------------------------------------------------------------------------
def _to_device(tree):
    from .device import asdevice
    from .jagged import JaggedArray
    from .table import Table
    if isinstance(tree, numpy.ndarray):
        return asdevice(tree)
    if isinstance(tree, HostJagged):
        content = _to_device(tree.content)
        if tree.counts is not None:
            return JaggedArray.fromcounts(asdevice(tree.counts), content)          # the offsets scan runs on the device
        return JaggedArray(asdevice(tree.starts), asdevice(tree.stops), content)
    if isinstance(tree, HostTable):
        if tree.rowstart != 0:
            raise NotImplementedError("Table row offsets are outside the GPU hot path")
        out = Table()
        for n, c in tree.columns.items():
            out[n] = _to_device(c)
        return out
    raise TypeError(type(tree))


def _to_host(x):
    from .device import DeviceArray
    from .jagged import JaggedArray
    from .masked import MaskedArray
    from .table import Table
    if isinstance(x, numpy.ndarray):
        return x
    if isinstance(x, DeviceArray):
        return x.get()
    if isinstance(x, JaggedArray):
        x._valid()
        if isinstance(x._content, MaskedArray):
            raise NotImplementedError("masked content (the result of pad) is not written; call fillna() first")
        if x._off64 is not None and len(x) > 0:                                     # jagged.py:290-295
            content = x._dense_content()
            if isinstance(content, JaggedArray) and not (content._off64 is not None):
                content = content.compact()
            return HostJagged(_to_host(content), counts=x.counts.get())
        return HostJagged(_to_host(x._content), starts=x.starts.get(), stops=x.stops.get())   # jagged.py:296-302
    if isinstance(x, Table):
        return HostTable([(n, _to_host(x[n])) for n in x.columns])
    raise TypeError("cannot write {0} to an .awkd file".format(type(x)))


def save(file, array, name=None, mode="a", **options):
    """awkward0.save (persist.py:661-717): one array, or a dict of named arrays, into a `.awkd` zip file."""
    arrays = array if isinstance(array, dict) else {"": array}
    if name is not None:
        arrays = dict((name + n, x) for n, x in arrays.items())
    save_host(file, OrderedDict((n, _to_host(x)) for n, x in arrays.items()), mode=mode, **options)


class Load(object):
    """awkward0.load's mapping (persist.py:720-767): arrays are read, inflated and moved to the device on access."""

    def __init__(self, file):
        self._file = zipfile.ZipFile(_path(file), mode="r")

    def __getitem__(self, where):
        return _to_device(read_tree(_ZipReader(self._file), where))

    def __iter__(self):
        for n in self._file.namelist():
            if n.endswith(".json"):
                yield n[:-5]

    def __len__(self):
        return sum(1 for _ in self)

    def keys(self):
        return list(self)

    def __contains__(self, where):
        return where in list(self)

    def __repr__(self):
        return "<awkward0_b200.load ({0} members)>".format(len(self))

    def close(self):
        self._file.close()

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()


def load(file):
    """awkward0.load (persist.py:719-727): the array itself when the file holds one unnamed array, else a mapping."""
    f = Load(file)
    if list(f) == [""]:
        out = f[""]
        f.close()
        return out
    return f
