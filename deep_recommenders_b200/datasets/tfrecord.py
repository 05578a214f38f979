"""TFRecord / tf.train.Example I/O without TensorFlow.

Reading is native (csrc/tfrecord.cu: dr_tfrecord_index, dr_example_parse_feature): the file is memory-mapped, the
record index is built once with both CRC-32C checks, and a batch of records is parsed column by column straight
into numpy buffers (int64 values / packed strings + offsets / row_splits) -- the replacement of
``tf.data.TFRecordDataset(...).batch(B).map(tf.io.parse_example)`` (reference datasets/movielens.py:114-131).

Writing (the reference's `_serialize_example` + `tf.io.TFRecordWriter`, datasets/movielens.py:54-62,78,95) is a
small pure-Python protobuf encoder; it exists so MovieLens can be converted here and so the tests can build files.
"""
from __future__ import annotations

import mmap
import os
import struct
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np

from .. import _lib
from ..feature_column import PackedStrings, RaggedFeature

INT64, BYTES, FLOAT = 0, 1, 2
_KINDS = {"int64": INT64, "string": BYTES, "bytes": BYTES, "float32": FLOAT, "float": FLOAT}


# ---- writer ------------------------------------------------------------------------------------------------------
def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:           # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _feature(value) -> bytes:
    """One tf.train.Feature from a python value: int / [int] -> Int64List (packed), float / [float] -> FloatList
    (packed), bytes / str / [bytes] -> BytesList -- the three list kinds `_serialize_example` uses."""
    vals = list(value) if isinstance(value, (list, tuple, np.ndarray)) else [value]
    if vals and all(isinstance(v, (bytes, str, np.bytes_, np.str_)) for v in vals):
        body = b"".join(_ld(1, v.encode("utf-8") if isinstance(v, str) else bytes(v)) for v in vals)
        return _ld(1, body)
    if vals and all(isinstance(v, (float, np.floating)) for v in vals):
        return _ld(2, _ld(1, struct.pack(f"<{len(vals)}f", *vals)))
    if all(isinstance(v, (int, np.integer)) for v in vals):
        return _ld(3, _ld(1, b"".join(_varint(int(v)) for v in vals)) if vals else b"")
    raise TypeError(f"cannot encode feature value {value!r}")


def serialize_example(features: Dict[str, object]) -> bytes:
    """tf.train.Example(features=Features(feature={...})).SerializeToString() (map entries in key order, as the
    C++ protobuf serializer emits them deterministically)."""
    entries = b"".join(_ld(1, _ld(1, k.encode("utf-8")) + _ld(2, _feature(v))) for k, v in sorted(features.items()))
    return _ld(1, entries)


class TFRecordWriter:
    """tf.io.TFRecordWriter: length | masked crc32c(length) | data | masked crc32c(data)."""

    def __init__(self, path: str):
        self._f = open(path, "wb")
        self._lib = _lib.load()

    def _mcrc(self, b: bytes) -> bytes:
        return struct.pack("<I", self._lib.dr_masked_crc32c_host(b, len(b)))

    def write(self, record: bytes) -> None:
        head = struct.pack("<Q", len(record))
        self._f.write(head + self._mcrc(head) + record + self._mcrc(record))

    def close(self) -> None:
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ---- reader ------------------------------------------------------------------------------------------------------
class TFRecordFile:
    """A memory-mapped TFRecord file with a record index; `parse(lo, hi, spec)` returns columnar features.

    spec: {name: ("int64" | "string" | "float32", fixed: bool)} -- fixed=True is tf.io.FixedLenFeature([], dtype)
    (exactly one value per record, anything else raises like TF does), fixed=False is tf.io.VarLenFeature(dtype).
    Fixed int64 / float -> numpy array [n]; fixed string -> PackedStrings; var-len -> RaggedFeature(values, row_splits)
    with values a numpy array or PackedStrings.
    """

    def __init__(self, path: str, verify_crc: bool = True):
        self._lib = _lib.load()
        self._fh = open(path, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ) if size else None
        self._buf = np.frombuffer(self._mm, dtype=np.uint8) if size else np.zeros(0, dtype=np.uint8)
        ptr = self._buf.ctypes.data if size else None
        n = self._lib.dr_tfrecord_index(ptr, size, int(verify_crc), None, None, 0)
        if n < 0:
            raise ValueError(f"{path}: {self._lib.dr_last_error().decode()}")
        self.offsets = np.empty(n, dtype=np.int64)
        self.lengths = np.empty(n, dtype=np.int64)
        if n:
            self._lib.dr_tfrecord_index(ptr, size, 0, self.offsets.ctypes.data, self.lengths.ctypes.data, n)
        self.path = path

    def __len__(self) -> int:
        return int(self.offsets.size)

    def record(self, i: int) -> bytes:
        o, n = int(self.offsets[i]), int(self.lengths[i])
        return bytes(self._buf[o:o + n])

    def _column(self, off, ln, name: str, kind: int):
        n = off.size
        splits = np.empty(n + 1, dtype=np.int64)
        tv, tb = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64)
        args = (self._buf.ctypes.data, off.ctypes.data, ln.ctypes.data, n, name.encode("utf-8"), kind)
        _lib.check(self._lib.dr_example_parse_feature(*args, splits.ctypes.data, None, None, None, tv.ctypes.data,
                                                      tb.ctypes.data), "dr_example_parse_feature")
        nv, nb = int(tv[0]), int(tb[0])
        if kind == BYTES:
            data = np.empty(max(nb, 1), dtype=np.uint8)
            voff = np.empty(nv + 1, dtype=np.int64)
            _lib.check(self._lib.dr_example_parse_feature(*args, None, None, data.ctypes.data, voff.ctypes.data, None, None),
                       "dr_example_parse_feature")
            return splits, PackedStrings(data[:nb] if nb else data[:0], voff)
        vals = np.empty(max(nv, 1), dtype=np.int64 if kind == INT64 else np.float32)
        _lib.check(self._lib.dr_example_parse_feature(*args, None, vals.ctypes.data, None, None, None, None),
                   "dr_example_parse_feature")
        return splits, vals[:nv]

    def _columns(self, off, ln, items):
        """All requested features in one native walk per record (dr_example_parse_batch), sizing pass then fill."""
        import ctypes as C
        n, nf = off.size, len(items)
        names = (C.c_char_p * nf)(*[name.encode("utf-8") for name, _ in items])
        kinds = (C.c_int * nf)(*[k for _, k in items])
        splits = [np.empty(n + 1, dtype=np.int64) for _ in items]
        tv, tb = np.zeros(nf, dtype=np.int64), np.zeros(nf, dtype=np.int64)
        ptrs = lambda arrs: (C.c_void_p * nf)(*[None if a is None else a.ctypes.data for a in arrs])
        head = (self._buf.ctypes.data, off.ctypes.data, ln.ctypes.data, n, nf, names, kinds, ptrs(splits))
        _lib.check(self._lib.dr_example_parse_batch(*head, None, None, None, tv.ctypes.data, tb.ctypes.data),
                   "dr_example_parse_batch")
        vals, byts, voffs = [], [], []
        for (_, k), nv, nb in zip(items, tv.tolist(), tb.tolist()):
            if k == BYTES:
                vals.append(None)
                byts.append(np.empty(max(nb, 1), dtype=np.uint8))
                voffs.append(np.empty(nv + 1, dtype=np.int64))
            else:
                vals.append(np.empty(max(nv, 1), dtype=np.int64 if k == INT64 else np.float32))
                byts.append(None)
                voffs.append(None)
        _lib.check(self._lib.dr_example_parse_batch(*head, ptrs(vals), ptrs(byts), ptrs(voffs), tv.ctypes.data,
                                                    tb.ctypes.data), "dr_example_parse_batch")
        out = []
        for (_, k), sp, v, b, vo, nv, nb in zip(items, splits, vals, byts, voffs, tv.tolist(), tb.tolist()):
            out.append((sp, PackedStrings(b[:nb], vo) if k == BYTES else v[:nv]))
        return out

    def parse(self, lo: int, hi: int, spec: Dict[str, Tuple[str, bool]]) -> Dict[str, object]:
        hi = min(hi, len(self))
        off = np.ascontiguousarray(self.offsets[lo:hi])
        ln = np.ascontiguousarray(self.lengths[lo:hi])
        for name, (dtype, _) in spec.items():
            if dtype not in _KINDS:
                raise ValueError(f"feature {name!r}: unsupported dtype {dtype!r}")
        items = [(name, _KINDS[dtype]) for name, (dtype, _) in spec.items()]
        out = {}
        for (name, (dtype, fixed)), (splits, vals) in zip(spec.items(), self._columns(off, ln, items) if items else []):
            if fixed:
                if not np.array_equal(splits, np.arange(off.size + 1)):
                    bad = int(np.flatnonzero(np.diff(splits) != 1)[0])
                    raise ValueError(f"Feature: {name} (data type: {dtype}) is required but could not be found / is not "
                                     f"a single value in record {lo + bad}")
                out[name] = vals
            else:
                out[name] = RaggedFeature(vals, splits)
        return out

    def close(self) -> None:
        self._buf = None
        if self._mm is not None:
            try:
                self._mm.close()
            except BufferError:
                pass
        self._fh.close()
