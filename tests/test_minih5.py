"""minih5: the pure-Python HDF5 subset the output files are written with (no HDF5 library in the image).

PARITY UNPINNED against libhdf5 (see the module docstring): these tests pin the writer against its own
independent reader and against structural invariants of the HDF5 file-format specification."""
import struct

import numpy as np
import pytest

from classicalspinmc.jl_b200 import minih5


def _sample():
    rng = np.random.default_rng(0)
    data = {
        "spins": rng.normal(size=(12, 3)),
        "lattice/size": np.array([4, 3], dtype=np.int64),
        "lattice/S": 0.5,
        "lattice/bc": "periodic",
        "unit_cell/bilinear/(1,2),(0, -1)": rng.normal(size=(3, 3)),
        "unit_cell/bilinear/(1,1),(1, 0)": np.eye(3),
        "unit_cell/quartic/(1,1,1,1),(1, 0),(0, 1),(1, 1)": rng.normal(size=(3, 3, 3, 3)),
        "f32": np.arange(5, dtype=np.float32),
        "i32": np.arange(-3, 3, dtype=np.int32).reshape(2, 3),
        "u8": np.array([0, 255], dtype=np.uint8),
        "flags": np.array([True, False, True]),
        "names": np.array(["ab", "c", "defg"]),
        "empty": np.zeros((0, 3)),
    }
    attrs = {"T": 0.3, "t_thermalization": 100000, "paramsfile": "/tmp/out/configuration.h5.params",
             "report": True, "vec": np.array([1.0, 2.0, 3.0]), "unicode": "µ=1"}
    return data, attrs


def _same(a, b):
    if isinstance(a, str) or isinstance(b, str):
        return a == b
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind in "SU" or b.dtype.kind in "SU":
        return a.shape == b.shape and [x.decode() if isinstance(x, bytes) else x for x in a.ravel().tolist()] == \
            [x.decode() if isinstance(x, bytes) else x for x in b.ravel().tolist()]
    return a.shape == b.shape and np.array_equal(a.astype(np.float64), b.astype(np.float64))


def test_roundtrip_all_supported_types(tmp_path):
    data, attrs = _sample()
    fn = str(tmp_path / "t.h5")
    minih5.write_file(fn, data, attrs, groups=["unit_cell/cubic", "observables"])
    d2, a2, g2 = minih5.read_file(fn)
    assert set(d2) == set(data) and set(a2) == set(attrs)
    for k in data:
        assert _same(d2[k], data[k]), k
    for k in attrs:
        assert _same(a2[k], attrs[k]), k
    assert d2["f32"].dtype == np.float32 and d2["i32"].dtype == np.int32 and d2["u8"].dtype == np.uint8
    assert d2["lattice/size"].dtype == np.int64 and isinstance(d2["lattice/bc"], str)
    # groups exist even when empty (src/hdf5.jl:64,71: create_group for cubic / quartic without terms)
    assert {"unit_cell", "unit_cell/bilinear", "unit_cell/cubic", "unit_cell/quartic", "lattice", "observables"} <= g2


def test_structure_follows_the_format_specification(tmp_path):
    data, attrs = _sample()
    fn = str(tmp_path / "t.h5")
    minih5.write_file(fn, data, attrs)
    raw = open(fn, "rb").read()
    # superblock version 0: signature, versions, 8-byte offsets/lengths, base 0, EOF == file size
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8:13] == b"\0" * 5 and raw[13] == 8 and raw[14] == 8
    leaf_k, internal_k = struct.unpack_from("<HH", raw, 16)
    base, free, eof, driver = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and free == driver == minih5.UNDEF and eof == len(raw) and len(raw) % 8 == 0
    name_off, root, cache, _ = struct.unpack_from("<QQII", raw, 56)
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert name_off == 0 and cache == 1 and root % 8 == 0
    # root object header: version 1, message block starts 16 bytes in, every message 8-aligned
    ver, _, nmsg, refcount, size = struct.unpack_from("<BBHII", raw, root)
    assert ver == 1 and refcount == 1 and size % 8 == 0 and nmsg == 1 + len(attrs)
    pos, seen = root + 16, []
    while pos < root + 16 + size:
        mtype, msize, _ = struct.unpack_from("<HHB", raw, pos)
        assert msize % 8 == 0
        seen.append(mtype)
        pos += 8 + msize
    assert pos == root + 16 + size and seen[0] == minih5.MSG_SYMBOL_TABLE and seen.count(minih5.MSG_ATTRIBUTE) == len(attrs)
    assert struct.unpack_from("<QQ", raw, root + 24) == (btree, heap)
    # group B-tree node and symbol-table node: signatures, full-size allocation, names sorted
    assert raw[btree:btree + 4] == b"TREE" and raw[btree + 4] == 0 and raw[btree + 5] == 0
    used, left, right = struct.unpack_from("<HQQ", raw, btree + 6)
    assert used == 1 and left == right == minih5.UNDEF
    key0, snod, key1 = struct.unpack_from("<QQQ", raw, btree + 24)
    assert key0 == 0 and raw[snod:snod + 4] == b"SNOD" and raw[snod + 4] == 1
    assert btree + 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8 <= len(raw)
    assert snod + 8 + 2 * leaf_k * 40 <= len(raw)
    assert raw[heap:heap + 4] == b"HEAP"
    seg_size, free_at, seg = struct.unpack_from("<QQQ", raw, heap + 8)
    assert seg % 8 == 0 and seg_size % 8 == 0 and free_at + 16 <= seg_size
    assert struct.unpack_from("<QQ", raw, seg + free_at) == (minih5.HEAP_FREE_NULL, seg_size - free_at)
    n = struct.unpack_from("<H", raw, snod + 6)[0]
    names = []
    for k in range(n):
        off, hdr = struct.unpack_from("<QQ", raw, snod + 8 + 40 * k)
        end = raw.index(b"\0", seg + off)
        names.append(raw[seg + off:end])
        assert hdr % 8 == 0 and off % 8 == 0
    assert names == sorted(names) and len(names) == len({p.split("/")[0] for p in data})
    end = raw.index(b"\0", seg + key1)
    assert raw[seg + key1:end] == names[-1]
    # IEEE double datatype message exactly as libhdf5 encodes H5T_IEEE_F64LE
    assert minih5._datatype_message(np.float64) == bytes.fromhex("11203f0008000000" "00004000340b0034ff030000")
    assert minih5._datatype_message(np.int64) == bytes.fromhex("1008000008000000" "00004000")


def test_file_object_modes_and_rewrite(tmp_path):
    fn = str(tmp_path / "c.h5")
    f = minih5.File(fn, "w")
    f.data["spins"] = np.ones((4, 3))
    f.data["@attrs/T"] = 0.25
    f.data["@attrs/paramsfile"] = "x.h5.params"
    f.close()
    size0 = len(open(fn, "rb").read())
    for it in range(3):                       # r+ rewrites do not grow the file (string sizes are stable)
        g = minih5.File(fn, "r+")
        g.data["spins"] = np.full((4, 3), float(it))
        g.close()
    assert len(open(fn, "rb").read()) == size0
    h = minih5.File(fn, "r")
    assert np.all(h.data["spins"] == 2.0) and h.data["@attrs/T"] == 0.25 and h.data["@attrs/paramsfile"] == "x.h5.params"
    h.close()
    with pytest.raises(FileNotFoundError):
        minih5.File(str(tmp_path / "missing.h5"), "r")


def test_rejects_what_it_does_not_understand(tmp_path):
    fn = str(tmp_path / "bad.h5")
    open(fn, "wb").write(b"not hdf5 at all" * 10)
    with pytest.raises(minih5.H5FormatError, match="signature"):
        minih5.read_file(fn)
    raw = bytearray(96)
    raw[:8] = minih5.SIGNATURE
    raw[8] = 2                                 # new-style superblock
    open(fn, "wb").write(raw)
    with pytest.raises(minih5.H5FormatError, match="superblock version 2"):
        minih5.read_file(fn)
    with pytest.raises(minih5.H5FormatError):
        minih5.write_file(fn, {"z": np.array([1 + 2j])})


def test_large_group_and_long_names(tmp_path):
    data = {f"g/key_{i:04d}_" + "x" * (i % 17): np.array([float(i)]) for i in range(300)}
    fn = str(tmp_path / "big.h5")
    minih5.write_file(fn, data)
    d2, _, groups = minih5.read_file(fn)
    assert set(d2) == set(data) and all(d2[k][0] == data[k][0] for k in data) and groups == {"g"}
