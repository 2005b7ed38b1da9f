"""Minimal pure-Python HDF5 writer / reader for the reference's output files (src/hdf5.jl).

There is no HDF5 library in the build image (no libhdf5, h5py, HDF5.jl), so the files the drivers write
(``*.h5.params``, ``configuration_<slot>.h5``, ``IC_*.h5``) are produced here directly from the published
HDF5 file-format specification, restricted to the subset HDF5 1.8-compatible writers emit by default and
every libhdf5 reads:

* superblock version 0, 8-byte offsets and lengths;
* version-1 object headers;
* old-style groups: symbol-table message -> version-1 B-tree (one leaf level) -> symbol-table nodes,
  names in a local heap;
* contiguous datasets (dataspace message v1, data-layout message v3, fill-value message v2) of
  little-endian IEEE floats, two's-complement integers and fixed-length strings;
* version-1 attribute messages on the root group (the reference only uses file-level attributes,
  src/hdf5.jl:6-31,166-167).

The reader handles the same subset plus what stock writers add around it (header continuation blocks,
compact layout, multi-level group B-trees, version-2 dataspaces, variable-length strings through the
global heap) and refuses anything else (new-style "OHDR" headers, chunked / filtered datasets) with an
explicit error instead of guessing.

PARITY UNPINNED: with no HDF5 implementation available the byte layout is checked only against this
module's own independent reader and against structural invariants of the specification
(tests/test_host_mirror.py); it has not been opened with libhdf5.

Array convention: datasets are stored row-major with the numpy shape (what h5py shows); a Julia array
written by HDF5.jl appears here with its dimensions reversed, exactly as in h5py (util/load.py:88-93).
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
HEAP_FREE_NULL = 1          # end of the local-heap free list on disk
GROUP_INTERNAL_K = 16

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8
MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE = 0xC, 0x10, 0x11


class H5FormatError(ValueError):
    pass


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- datatype / dataspace messages ------------------------------------------------------------------------
def _datatype_message(dt: np.dtype) -> bytes:
    """Datatype message body (version 1) for a numpy dtype of the supported subset."""
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise H5FormatError("big-endian data is not supported")
    if dt.kind == "f":
        if dt.itemsize == 8:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        elif dt.itemsize == 4:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        else:
            raise H5FormatError(f"unsupported float size {dt.itemsize}")
        # class 1, bit field: little-endian, mantissa normalisation 2 (msb implied), sign bit position
        head = struct.pack("<BBBBI", 0x11, 0x20, sign, 0x00, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esize, 0, msize, bias)
    if dt.kind in "iub":
        signed = 0x08 if dt.kind == "i" else 0x00
        head = struct.pack("<BBBBI", 0x10, signed, 0x00, 0x00, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        # class 3: null-terminated (0), character set UTF-8 (1 << 4): ASCII is a subset
        return struct.pack("<BBBBI", 0x13, 0x10, 0x00, 0x00, max(dt.itemsize, 1))
    raise H5FormatError(f"unsupported dtype {dt}")


def _dataspace_message(shape) -> bytes:
    """Dataspace message version 1 (simple dataspace, no maximum dimensions); rank 0 is a scalar."""
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _normalise(value):
    """numpy array of a storable dtype for a Python / numpy value (str -> fixed-length bytes)."""
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, (bytes, np.bytes_)):
        b = bytes(value)
        return np.array(b, dtype=f"S{len(b) + 1}")          # room for the terminating NUL
    a = np.asarray(value)
    if a.dtype.kind == "U":
        return _normalise(np.char.encode(a, "utf-8"))
    if a.dtype.kind == "S":
        longest = max([len(x) for x in a.ravel().tolist()] or [0])
        return a.astype(f"S{longest + 1}")                   # fixed length = longest string + NUL
    if a.dtype.kind == "b":
        return a.astype(np.int8)
    if a.dtype.kind == "f" and a.dtype.itemsize not in (4, 8):
        return a.astype(np.float64)
    if a.dtype.kind not in "fiu":
        raise H5FormatError(f"cannot store values of dtype {a.dtype}")
    return a.astype(a.dtype.newbyteorder("<")) if a.dtype.byteorder == ">" else a


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages) -> bytes:
    data = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


# ---- writer ------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self, leaf_k: int):
        self.leaf_k = leaf_k
        self.buf = bytearray(96)     # superblock goes here

    def put(self, b: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += b
        return addr

    def dataset(self, value) -> int:
        a = np.asarray(_normalise(value), order="C")
        raw = a.tobytes()
        addr = self.put(raw) if raw else UNDEF
        msgs = [
            _message(MSG_DATASPACE, _dataspace_message(a.shape)),
            _message(MSG_DATATYPE, _datatype_message(a.dtype), flags=1),
            _message(MSG_FILL, struct.pack("<BBBB", 2, 2, 2, 0), flags=1),   # late allocation, fill if set, none defined
            _message(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, addr, len(raw))),
        ]
        return self.put(_object_header(msgs))

    def group(self, children: dict, attrs: dict | None = None):
        """children: name -> ndarray-like (dataset) or dict (sub-group).  Returns (header, btree, heap)."""
        entries = []
        for name in sorted(children, key=lambda s: s.encode("utf-8")):
            child = children[name]
            if isinstance(child, dict):
                hdr, bt, hp = self.group(child)
                entries.append((name, hdr, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name, self.dataset(child), 0, b"\0" * 16))
        if len(entries) > 2 * self.leaf_k:
            raise H5FormatError("group larger than the symbol-table node size chosen for the file")
        # local heap: "" at offset 0, names 8-aligned, one free block at the end
        seg = bytearray(8)
        offsets = []
        for name, *_ in entries:
            offsets.append(len(seg))
            seg += _pad8(name.encode("utf-8") + b"\0")
        free_at = len(seg)
        seg += struct.pack("<QQ", HEAP_FREE_NULL, 32) + b"\0" * 16
        seg_addr = self.put(bytes(seg))
        heap = self.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_at, seg_addr))
        # one symbol-table node holding every entry, allocated at its full size
        if entries:
            snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(entries)))
            for (name, hdr, cache, scratch), off in zip(entries, offsets):
                snod += struct.pack("<QQII", off, hdr, cache, 0) + scratch
            snod += b"\0" * (8 + 2 * self.leaf_k * 40 - len(snod))
            snod_addr = self.put(bytes(snod))
        # B-tree node (type 0 = group, level 0): key0 = "", key1 = largest name in the only child
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, UNDEF, UNDEF))
        if entries:
            node += struct.pack("<QQQ", 0, snod_addr, offsets[-1])
        node += b"\0" * (24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8 - len(node))
        btree = self.put(bytes(node))
        msgs = [_message(MSG_SYMBOL_TABLE, struct.pack("<QQ", btree, heap))]
        for name in sorted(attrs or {}):
            a = np.asarray(_normalise(attrs[name]), order="C")
            nm = name.encode("utf-8") + b"\0"
            dt, ds = _datatype_message(a.dtype), _dataspace_message(a.shape)
            body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()
            if len(body) > 65000:
                raise H5FormatError(f"attribute {name!r} too large for a version-1 header message")
            msgs.append(_message(MSG_ATTRIBUTE, body))
        return self.put(_object_header(msgs)), btree, heap


def _tree_from_paths(data: dict, groups=()):
    root: dict = {}
    for g in groups:
        node = root
        for part in [p for p in g.split("/") if p]:
            node = node.setdefault(part, {})
    for path, value in data.items():
        parts = [p for p in path.split("/") if p]
        node = root
        for part in parts[:-1]:
            node = node.setdefault(part, {})
            if not isinstance(node, dict):
                raise H5FormatError(f"{path!r}: a dataset is in the way")
        node[parts[-1]] = value
    return root


def _max_group_size(tree: dict) -> int:
    return max([len(tree)] + [_max_group_size(v) for v in tree.values() if isinstance(v, dict)])


def write_file(filename, data: dict, attrs: dict | None = None, groups=()):
    """data: {"group/sub/name": array-like}; attrs: root attributes; groups: paths of (possibly empty) groups."""
    tree = _tree_from_paths(data, groups)
    leaf_k = max(4, (_max_group_size(tree) + 1) // 2)
    if leaf_k > 0x7FFF:
        raise H5FormatError("too many entries in one group")
    w = _Writer(leaf_k)
    root, btree, heap = w.group(tree, attrs)
    eof = len(w.buf) + (-len(w.buf) % 8)
    w.buf += b"\0" * (eof - len(w.buf))
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, GROUP_INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)     # root symbol-table entry
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(filename, "wb") as f:
        f.write(w.buf)


# ---- reader ------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, raw: bytes):
        self.raw = raw
        if raw[:8] != SIGNATURE:
            raise H5FormatError("not an HDF5 file (signature at offset 0 missing)")
        ver = raw[8]
        if ver not in (0, 1):
            raise H5FormatError(f"superblock version {ver} (new-style file) is not supported by this reader")
        if raw[13] != 8 or raw[14] != 8:
            raise H5FormatError("only 8-byte offsets and lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", raw, off)
        self.root_header = struct.unpack_from("<QQ", raw, off + 32)[1]

    # -- object headers -----------------------------------------------------------------------------------
    def messages(self, addr: int):
        raw = self.raw
        if raw[addr:addr + 4] == b"OHDR":
            raise H5FormatError("version-2 object headers are not supported by this reader")
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", raw, addr)
        if ver != 1:
            raise H5FormatError(f"object header version {ver} at {addr}")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", raw, pos)
                body = raw[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == MSG_CONTINUATION:
                    blocks.append(struct.unpack_from("<QQ", body))
                out.append((mtype, body))
        return out

    # -- groups ----------------------------------------------------------------------------------------------
    def heap_name(self, heap_addr: int, offset: int) -> str:
        raw = self.raw
        if raw[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("local heap signature missing")
        seg = struct.unpack_from("<Q", raw, heap_addr + 24)[0]
        end = raw.index(b"\0", seg + offset)
        return raw[seg + offset:end].decode("utf-8")

    def group_entries(self, btree: int, heap: int):
        raw = self.raw
        if raw[btree:btree + 4] != b"TREE":
            raise H5FormatError("group B-tree signature missing")
        ntype, level, used = struct.unpack_from("<BBH", raw, btree + 4)
        if ntype != 0:
            raise H5FormatError("not a group B-tree")
        out = []
        for i in range(used):
            child = struct.unpack_from("<Q", raw, btree + 24 + 8 + 16 * i)[0]
            if level > 0:
                out += self.group_entries(child, heap)
                continue
            if raw[child:child + 4] != b"SNOD":
                raise H5FormatError("symbol-table node signature missing")
            n = struct.unpack_from("<H", raw, child + 6)[0]
            for k in range(n):
                name_off, hdr = struct.unpack_from("<QQ", raw, child + 8 + 40 * k)
                out.append((self.heap_name(heap, name_off), hdr))
        return out

    # -- datasets --------------------------------------------------------------------------------------------
    @staticmethod
    def dataspace(body: bytes):
        ver, rank, flags = struct.unpack_from("<BBB", body)
        if ver == 1:
            off = 8
        elif ver == 2:
            if body[3] == 2:
                return None     # null dataspace
            off = 4
        else:
            raise H5FormatError(f"dataspace version {ver}")
        return tuple(struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank))

    @staticmethod
    def datatype(body: bytes):
        """-> (numpy dtype or ("vlen-str",), element size)"""
        cv, b0, b1, _b2, size = struct.unpack_from("<BBBBI", body)
        cls = cv & 0x0F
        if cls in (0, 1) and (b0 & 1):
            raise H5FormatError("big-endian data is not supported")
        if cls == 0:
            return np.dtype(("<i" if b0 & 0x08 else "<u") + str(size)), size
        if cls == 1:
            if size not in (4, 8):
                raise H5FormatError(f"unsupported float size {size}")
            return np.dtype(f"<f{size}"), size
        if cls == 3:
            return np.dtype(f"S{size}"), size
        if cls == 9 and (b0 & 0x0F) == 1:
            return ("vlen-str",), size
        raise H5FormatError(f"unsupported datatype class {cls}")

    def global_heap_object(self, addr: int, index: int) -> bytes:
        raw = self.raw
        if raw[addr:addr + 4] != b"GCOL":
            raise H5FormatError("global heap signature missing")
        size = struct.unpack_from("<Q", raw, addr + 8)[0]
        pos = addr + 16
        while pos + 16 <= addr + size:
            idx, _, osize = struct.unpack_from("<HH4xQ", raw, pos)
            if idx == 0:
                break
            if idx == index:
                return raw[pos + 16:pos + 16 + osize]
            pos += 16 + osize + (-osize % 8)
        raise H5FormatError("global heap object not found")

    def decode(self, dt, shape, data: bytes):
        if shape is None:
            return None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if isinstance(dt, tuple):       # variable-length strings: (length u32, heap address u64, index u32) each
            vals = []
            for i in range(n):
                length, addr, idx = struct.unpack_from("<IQI", data, 16 * i)
                vals.append(self.global_heap_object(addr, idx)[:length].decode("utf-8") if length else "")
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        a = np.frombuffer(data[:n * dt.itemsize], dtype=dt).reshape(shape).copy()
        if dt.kind == "S":
            return a[()].decode("utf-8") if not shape else a
        return a[()] if not shape else a

    def dataset(self, msgs):
        shape = dt = None
        data = b""
        for mtype, body in msgs:
            if mtype == MSG_DATASPACE:
                shape = self.dataspace(body)
            elif mtype == MSG_DATATYPE:
                dt, _ = self.datatype(body)
            elif mtype == MSG_LAYOUT:
                ver, cls = body[0], body[1]
                if ver != 3:
                    raise H5FormatError(f"data layout version {ver} is not supported")
                if cls == 0:
                    size = struct.unpack_from("<H", body, 2)[0]
                    data = body[4:4 + size]
                elif cls == 1:
                    addr, size = struct.unpack_from("<QQ", body, 2)
                    data = b"" if addr == UNDEF else self.raw[self.base + addr:self.base + addr + size]
                else:
                    raise H5FormatError("chunked datasets are not supported by this reader")
        if dt is None:
            raise H5FormatError("dataset without a datatype message")
        return self.decode(dt, shape, data)

    def attribute(self, body: bytes):
        ver = body[0]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            pos = 8
            name = body[pos:pos + nsz].split(b"\0")[0].decode("utf-8"); pos += nsz + (-nsz % 8)
            tbody = body[pos:pos + tsz]; pos += tsz + (-tsz % 8)
            sbody = body[pos:pos + ssz]; pos += ssz + (-ssz % 8)
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            pos = 8 + (1 if ver == 3 else 0)
            name = body[pos:pos + nsz].split(b"\0")[0].decode("utf-8"); pos += nsz
            tbody = body[pos:pos + tsz]; pos += tsz
            sbody = body[pos:pos + ssz]; pos += ssz
        else:
            raise H5FormatError(f"attribute message version {ver}")
        dt, _ = self.datatype(tbody)
        return name, self.decode(dt, self.dataspace(sbody), body[pos:])

    # -- walk ------------------------------------------------------------------------------------------------
    def walk(self, header: int, prefix: str, data: dict, groups: set, attrs: dict | None):
        msgs = self.messages(header)
        sym = [b for t, b in msgs if t == MSG_SYMBOL_TABLE]
        if attrs is not None:
            for t, b in msgs:
                if t == MSG_ATTRIBUTE:
                    k, v = self.attribute(b)
                    attrs[k] = v
        if sym:
            if prefix:
                groups.add(prefix.rstrip("/"))
            btree, heap = struct.unpack_from("<QQ", sym[0])
            for name, hdr in self.group_entries(btree, heap):
                self.walk(hdr, prefix + name + "/", data, groups, None)
        elif any(t == MSG_LAYOUT for t, _ in msgs):
            data[prefix.rstrip("/")] = self.dataset(msgs)
        else:
            raise H5FormatError(f"object {prefix!r}: neither an old-style group nor a dataset")


def read_file(filename):
    """-> (data {path: value}, root attributes {name: value}, set of group paths)"""
    with open(filename, "rb") as f:
        raw = f.read()
    r = _Reader(raw)
    data, attrs, groups = {}, {}, set()
    r.walk(r.root_header, "", data, groups, attrs)
    return data, attrs, groups


class File:
    """Whole-file container with the interface hdf5.py needs: ``data`` (path -> value; root attributes under
    ``@attrs/<name>``), ``groups`` (paths of groups that exist even when empty); modes "w", "r", "r+".  The
    file is rewritten in full on ``close()`` — the reference's files are a few datasets each."""

    def __init__(self, filename, mode):
        self.filename, self.mode = filename, mode
        self.data, self.groups = {}, set()
        if mode in ("r", "r+"):
            data, attrs, groups = read_file(filename)
            self.data = dict(data)
            self.data.update({"@attrs/" + k: v for k, v in attrs.items()})
            self.groups = set(groups)
        elif mode != "w":
            raise ValueError(f"mode {mode!r}")

    def close(self):
        if self.mode == "r":
            return
        data = {k: v for k, v in self.data.items() if not k.startswith("@attrs/")}
        attrs = {k[len("@attrs/"):]: v for k, v in self.data.items() if k.startswith("@attrs/")}
        write_file(self.filename, data, attrs, self.groups)
