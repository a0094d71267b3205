"""TensorFlow V2 checkpoint ("tensor bundle") reader -- and a minimal writer -- without TensorFlow.

The reference restores its model with ``tf.train.Saver().restore(sess, options.model_path)``
(``src/network/predict.py:181-184``; ``-m`` flag ``SVision:35``; the README asks for the three files
``svision-cnn-model.ckpt.{index,data-00000-of-00001,meta}`` and the ``...ckpt`` prefix,
``README.md:68,85-86``).  TensorFlow is not installable in this image, so this module parses the
on-disk format directly (public format: tensorflow/core/util/tensor_bundle, leveldb table):

* ``<prefix>.index`` is a LevelDB-format table: 48-byte footer (metaindex handle, index handle,
  magic ``0xdb4775248b80fb57``), prefix-compressed key/value blocks each followed by a 1-byte
  compression tag and a masked crc32c.  Key ``""`` holds ``BundleHeaderProto``; every other key is
  a variable name whose value is a ``BundleEntryProto`` (1 dtype, 2 shape, 3 shard_id, 4 offset,
  5 size, 6 crc32c fixed32, 7 slices);
* tensor bytes sit raw, little-endian, in ``<prefix>.data-%05d-of-%05d``.

Status: **round-trip tested only** (writer <-> reader, plus structural checks against the format
description).  No real ``svision-cnn-model.ckpt`` exists in this environment (it is a Google-Drive
download), so parity on the real checkpoint is unpinned; the reader fails loudly on anything it
does not understand (compressed blocks, sliced tensors, non-float dtypes, crc mismatches).
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT = 1
_MASK_DELTA = 0xA282EAD8

# ---- crc32c (Castagnoli), table driven; numpy-vectorised slicing for large buffers ----------------
_POLY = 0x82F63B78
_T = np.zeros((8, 256), dtype=np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _T[0, _i] = _c
for _k in range(1, 8):
    _T[_k] = (_T[_k - 1] >> 8) ^ _T[0][_T[_k - 1] & 0xFF]
_T0 = [int(v) for v in _T[0]]


def crc32c(data: bytes, crc: int = 0) -> int:
    crc ^= 0xFFFFFFFF
    mv = memoryview(data)
    n8 = len(mv) // 8 * 8
    if n8 >= 64:
        # slicing-by-8, eight table lookups per 8 bytes; the loop over words stays in Python but
        # touches 8 bytes per iteration via a uint32 view (fast enough for index blocks; tensor
        # payloads are verified with verify_data=True only on request)
        w = np.frombuffer(mv[:n8], dtype="<u4").astype(np.uint64)
        t = [[int(x) for x in _T[k]] for k in range(8)]
        for i in range(0, w.size, 2):
            lo = int(w[i]) ^ crc
            hi = int(w[i + 1])
            crc = (t[7][lo & 0xFF] ^ t[6][(lo >> 8) & 0xFF] ^ t[5][(lo >> 16) & 0xFF] ^ t[4][lo >> 24] ^
                   t[3][hi & 0xFF] ^ t[2][(hi >> 8) & 0xFF] ^ t[1][(hi >> 16) & 0xFF] ^ t[0][hi >> 24])
        mv = mv[n8:]
    for b in mv:
        crc = _T0[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ---- varints / protobuf wire helpers ----------------------------------------------------------------
def _get_varint(buf: bytes, pos: int):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf: bytes) -> dict:
    """field number -> list of raw values (ints for varint/fixed, bytes for length-delimited)."""
    out = {}
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _parse_shape(buf: bytes):
    dims = []
    for d in _parse_proto(buf).get(2, []):          # TensorShapeProto.dim
        size = _parse_proto(d).get(1, [0])[0]       # Dim.size (int64 varint)
        if size >= 1 << 63:
            size -= 1 << 64
        dims.append(int(size))
    return tuple(dims)


# ---- leveldb table -----------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    block = data[offset:offset + size]
    tag = data[offset + size]
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if tag != 0:
        raise ValueError(f"checkpoint index block is compressed (tag {tag}); only uncompressed "
                         "tables, as TensorFlow writes them, are supported")
    if verify and mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
        raise ValueError("checkpoint index block failed its crc32c check")
    return block


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(prefix: str, verify: bool = True) -> dict:
    """``{name: {'dtype', 'shape', 'shard_id', 'offset', 'size', 'crc32c'}}`` plus ``''`` -> header."""
    path = prefix + ".index"
    data = open(path, "rb").read()
    if len(data) < 48:
        raise ValueError(f"{path}: too short to be a checkpoint index")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: bad table magic (not a TensorFlow V2 checkpoint index)")
    pos = 0
    _, pos = _get_varint(footer, pos)               # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    entries = {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        for key, value in _block_entries(_read_block(data, off, size, verify)):
            name = key.decode()
            msg = _parse_proto(value)
            if name == "":
                entries[""] = {"num_shards": msg.get(1, [1])[0], "endianness": msg.get(2, [0])[0]}
                continue
            if 7 in msg:
                raise ValueError(f"variable {name!r} is stored in slices; not supported")
            entries[name] = {"dtype": msg.get(1, [0])[0], "shape": _parse_shape(msg.get(2, [b""])[0]),
                             "shard_id": msg.get(3, [0])[0], "offset": msg.get(4, [0])[0],
                             "size": msg.get(5, [0])[0], "crc32c": msg.get(6, [None])[0]}
    if "" in entries and entries[""]["endianness"] != 0:
        raise ValueError("big-endian checkpoints are not supported")
    return entries


def read_bundle(prefix: str, names=None, verify_data: bool = False) -> dict:
    """``{name: float32 ndarray}`` for ``names`` (default: every float variable in the bundle)."""
    index = read_index(prefix)
    num_shards = index.get("", {}).get("num_shards", 1)
    wanted = [n for n in index if n != ""] if names is None else list(names)
    out = {}
    shards = {}
    for name in wanted:
        if name not in index:
            raise KeyError(f"checkpoint {prefix!r} has no variable {name!r}; it holds "
                           f"{sorted(n for n in index if n)[:20]}")
        e = index[name]
        if e["dtype"] != DT_FLOAT:
            if names is None:
                continue
            raise TypeError(f"variable {name!r} has dtype enum {e['dtype']}, expected DT_FLOAT (1)")
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        if e["size"] != 4 * count:
            raise ValueError(f"variable {name!r}: {e['size']} bytes for shape {e['shape']}")
        if verify_data and e["crc32c"] is not None and mask_crc(crc32c(bytes(raw))) != e["crc32c"]:
            raise ValueError(f"variable {name!r} failed its crc32c check")
        out[name] = np.frombuffer(bytes(raw), dtype="<f4").reshape(e["shape"]).copy()
    return out


# ---- writer (tests / exporting synthetic weights in the reference's own format) ----------------------
def _entry_proto(arr: np.ndarray, offset: int, crc: int) -> bytes:
    shape = b"".join(b"\x12" + _put_varint(len(d)) + d
                     for d in (b"\x08" + _put_varint(int(s)) for s in arr.shape))
    return (b"\x08" + _put_varint(DT_FLOAT) + b"\x12" + _put_varint(len(shape)) + shape +
            b"\x20" + _put_varint(offset) + b"\x28" + _put_varint(arr.nbytes) +
            b"\x35" + struct.pack("<I", crc))


def _build_block(items) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        if i % 16 == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        out += key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix: str, tensors: dict, data_crc: bool = True) -> None:
    """Write ``{name: float32 array}`` as ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``."""
    names = sorted(tensors)
    items = [(b"", b"\x08\x01\x10\x00\x1a\x02\x08\x01")]        # header: 1 shard, little endian, version
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in names:
            arr = np.array(tensors[name], dtype="<f4", order="C")      # keeps 0-d scalars 0-d
            raw = arr.tobytes()
            f.write(raw)
            crc = mask_crc(crc32c(raw)) if data_crc else 0
            items.append((name.encode(), _entry_proto(arr, offset, crc)))
            offset += len(raw)
    out = bytearray()

    def emit(block: bytes):
        off = len(out)
        out.extend(block)
        out.append(0)
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return off, len(block)

    d_off, d_size = emit(_build_block(items))
    m_off, m_size = emit(_build_block([]))
    last_key = items[-1][0] + b"\x00"
    i_off, i_size = emit(_build_block([(last_key, _put_varint(d_off) + _put_varint(d_size))]))
    footer = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
