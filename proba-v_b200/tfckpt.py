"""TensorFlow checkpoint ("tensor bundle") reader / writer for the reference's tf.train.Checkpoint files.

The reference saves `tf.train.Checkpoint(step, psnr, optimizer, model)` through a `CheckpointManager(max_to_keep=5)`
(models/trainClass.py:33-39, restored at :52-59 and in test.py:58-67).  On disk that is

    <prefix>.index                    a leveldb-format sorted string table: key "" -> BundleHeaderProto, every other
                                      key -> BundleEntryProto {dtype, shape, shard_id, offset, size, masked crc32c}
    <prefix>.data-0000s-of-0000n      the raw little-endian tensor bytes of shard s
    checkpoint                        text-format CheckpointState (model_checkpoint_path, all_model_checkpoint_paths)

with variable keys `model/layer_with_weights-N/{v,g,layer/bias,initialized}/.ATTRIBUTES/VARIABLE_VALUE`, optimizer slots
`<var key minus suffix>/.OPTIMIZER_SLOT/optimizer/{m,v}/.ATTRIBUTES/VARIABLE_VALUE`, `optimizer/{iter,beta_1,beta_2,decay,
learning_rate,momentum_cache}`, `psnr`, `step`, `save_counter` and the serialized TrackableObjectGraph under
`_CHECKPOINTABLE_OBJECT_GRAPH` (parsed from modelInfo/ckpt_p16t9c85r12/NIR/ckpt-124.index).

TensorFlow is not installable in this environment, so the formats are implemented here from their published layouts
(leveldb table_format.md; tensorflow/core/protobuf/tensor_bundle.proto; trackable_object_graph.proto) and pinned by the
reference's own checkpoint files: tests/test_tfckpt.py parses the shipped `.index` + shard-0 files (committed as fixtures
under tests/golden/tf_ckpt/), checks every block / tensor crc32c, and requires a bundle written here for the same graph
to carry exactly the reference's key set, dtypes, shapes and object-graph topology.

Host-side code only (pure Python + numpy): the weights travel to / from the device through WDSRModel.set_weights /
get_weights and pv_trainer_{get,set}_state.
"""
from __future__ import annotations

import os
import struct
import time
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli)
_CRC_POLY = 0x82F63B78


def _make_tables():
    t0 = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (_CRC_POLY if c & 1 else 0)
        t0[i] = c
    tabs = [t0]
    for _ in range(7):
        prev = tabs[-1]
        tabs.append((prev >> 8) ^ t0[prev & 0xFF])
    return [t.tolist() for t in tabs]


_T = _make_tables()


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C of `data` (slicing-by-8, pure Python; ~15 MB/s, a 535 267-parameter checkpoint is 6.4 MB)."""
    c = crc ^ 0xFFFFFFFF
    mv = memoryview(data).cast("B")
    n8 = len(mv) // 8
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    if n8:
        for lo, hi in struct.iter_unpack("<II", mv[:n8 * 8]):
            lo ^= c
            c = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24] ^
                 t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    for b in mv[n8 * 8:]:
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:
    """leveldb / TF crc masking: rotate right by 15 and add a constant."""
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m: int) -> int:
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ protobuf wire format
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos: int) -> Tuple[int, int]:
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def pb_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field number, wire type, value)] of one message level; length-delimited values stay bytes."""
    out, pos = [], 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((f, wt, v))
    return out


def _pb_varint(f: int, v: int) -> bytes:
    return _varint(f << 3) + _varint(v)


def _pb_bytes(f: int, b: bytes) -> bytes:
    return _varint((f << 3) | 2) + _varint(len(b)) + b


def _pb_fixed32(f: int, v: int) -> bytes:
    return _varint((f << 3) | 5) + struct.pack("<I", v)


# DataType enum values (tensorflow/core/framework/types.proto) of the dtypes a Checkpoint of this model holds
DT_FLOAT, DT_INT32, DT_STRING, DT_INT64, DT_BOOL = 1, 3, 7, 9, 10
_NP_OF_DT = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8"), DT_BOOL: np.dtype("bool")}
_DT_OF_NP = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64, np.dtype("bool"): DT_BOOL}


class BundleEntry:
    """BundleEntryProto (tensor_bundle.proto): where a tensor's bytes live and their checksum."""

    __slots__ = ("dtype", "shape", "shard_id", "offset", "size", "crc32c")

    def __init__(self, dtype=0, shape=(), shard_id=0, offset=0, size=0, crc=0):
        self.dtype, self.shape, self.shard_id, self.offset, self.size, self.crc32c = dtype, tuple(shape), shard_id, offset, size, crc

    @classmethod
    def parse(cls, buf: bytes) -> "BundleEntry":
        e = cls()
        for f, wt, v in pb_fields(buf):
            if f == 1:
                e.dtype = v
            elif f == 2:                                    # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
                dims = []
                for f2, _, v2 in pb_fields(v):
                    if f2 == 2:
                        size = 0
                        for f3, _, v3 in pb_fields(v2):
                            if f3 == 1:
                                size = v3
                        dims.append(size)
                e.shape = tuple(dims)
            elif f == 3:
                e.shard_id = v
            elif f == 4:
                e.offset = v
            elif f == 5:
                e.size = v
            elif f == 6:
                e.crc32c = struct.unpack("<I", v)[0]
        return e

    def serialize(self) -> bytes:
        out = b""
        if self.dtype:
            out += _pb_varint(1, self.dtype)
        shape = b"".join(_pb_bytes(2, _pb_varint(1, d)) for d in self.shape)
        out += _pb_bytes(2, shape)
        if self.shard_id:
            out += _pb_varint(3, self.shard_id)
        if self.offset:
            out += _pb_varint(4, self.offset)
        if self.size:
            out += _pb_varint(5, self.size)
        out += _pb_fixed32(6, self.crc32c)
        return out


# ------------------------------------------------------------------------------------------------ leveldb table
_TABLE_MAGIC = 0xDB4775248B80FB57
_BLOCK_RESTART_INTERVAL = 16
_BLOCK_SIZE = 262144             # tensorflow::table::Options::block_size (the bundle index of this model is one 31 KB block)


def _read_block(buf: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    data = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack("<I", buf[offset + size + 1:offset + size + 5])[0]
    if verify and unmask_crc(stored) != crc32c(buf[offset:offset + size + 1]):
        raise ValueError(f"table block at {offset}: crc32c mismatch")
    if ctype != 0:
        raise ValueError(f"table block at {offset}: compression type {ctype} not supported (TF writes bundles uncompressed)")
    return data


def _block_entries(block: bytes) -> Iterable[Tuple[bytes, bytes]]:
    nrestart = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * nrestart
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a leveldb-format table file, in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _TABLE_MAGIC:
        raise ValueError(f"{path}: not a leveldb table (bad magic)")
    footer = buf[-48:]
    pos = 0
    _, pos = _read_varint(footer, pos)          # metaindex handle (unused by TF)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p = _read_varint(handle, 0)
        bsize, _ = _read_varint(handle, p)
        out.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
    return out


class _BlockBuilder:
    def __init__(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count % _BLOCK_RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def size(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start: bytes, limit: bytes) -> bytes:
    """leveldb BytewiseComparator::FindShortestSeparator: a short key in [start, limit)."""
    n = min(len(start), len(limit))
    d = 0
    while d < n and start[d] == limit[d]:
        d += 1
    if d < n and start[d] < 0xFF and start[d] + 1 < limit[d]:
        return start[:d] + bytes([start[d] + 1])
    return start


def _short_successor(key: bytes) -> bytes:
    """leveldb BytewiseComparator::FindShortSuccessor: a short key >= key."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_table(path: str, items: List[Tuple[bytes, bytes]]):
    """Writes sorted (key, value) pairs as an uncompressed leveldb table, byte for byte what tensorflow::table::TableBuilder
    emits (restart interval 16, index keys shortened with the bytewise comparator; tests/test_tfckpt.py re-writes the
    reference's own index file and compares the bytes)."""
    items = sorted(items, key=lambda kv: kv[0])
    out = bytearray()
    index = _BlockBuilder()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block)
        out.append(0)                                                     # kNoCompression
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _varint(off) + _varint(len(block))

    cur = _BlockBuilder()
    pending = None                                         # (last key, handle) of a flushed block waiting for its index key
    for k, v in items:
        if pending is not None:
            index.add(_shortest_separator(pending[0], k), pending[1])
            pending = None
        cur.add(k, v)
        if cur.size() >= _BLOCK_SIZE:
            pending = (cur.last, emit(cur.finish()))
            cur = _BlockBuilder()
    if cur.count:
        pending = (cur.last, emit(cur.finish()))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(out)


# ------------------------------------------------------------------------------------------------ tensor bundle
def _shard_name(prefix: str, shard: int, nshards: int) -> str:
    return f"{prefix}.data-{shard:05d}-of-{nshards:05d}"


class BundleReader:
    """tensorflow::BundleReader for the dtypes a tf.train.Checkpoint of this model contains."""

    def __init__(self, prefix: str, verify_table: bool = True):
        self.prefix = prefix
        self.entries: Dict[str, BundleEntry] = {}
        self.num_shards, self.endianness, self.version = 1, 0, (0, 0)
        for k, v in read_table(prefix + ".index", verify_table):
            if k == b"":
                for f, _, val in pb_fields(v):                    # BundleHeaderProto
                    if f == 1:
                        self.num_shards = val
                    elif f == 2:
                        self.endianness = val
                    elif f == 3:
                        vd = {ff: vv for ff, _, vv in pb_fields(val)}
                        self.version = (vd.get(1, 0), vd.get(2, 0))
            else:
                self.entries[k.decode()] = BundleEntry.parse(v)
        if self.endianness != 0:
            raise ValueError("big-endian bundles are not supported")
        self._shards: Dict[int, Optional[bytes]] = {}

    def keys(self) -> List[str]:
        return sorted(self.entries)

    def has_shard(self, shard: int) -> bool:
        return os.path.exists(_shard_name(self.prefix, shard, self.num_shards))

    def _shard(self, shard: int) -> bytes:
        if shard not in self._shards:
            p = _shard_name(self.prefix, shard, self.num_shards)
            if not os.path.exists(p):
                raise FileNotFoundError(f"{p}: data shard missing (the reference repository ships only shard 0 of its checkpoints)")
            self._shards[shard] = open(p, "rb").read()
        return self._shards[shard]

    def raw(self, key: str, verify: bool = True) -> bytes:
        e = self.entries[key]
        b = self._shard(e.shard_id)[e.offset:e.offset + e.size]
        if len(b) != e.size:
            raise ValueError(f"{key}: shard {e.shard_id} is truncated")
        if verify and e.dtype != DT_STRING and unmask_crc(e.crc32c) != crc32c(b):
            raise ValueError(f"{key}: tensor crc32c mismatch")
        return b

    def tensor(self, key: str, verify: bool = True):
        """numpy array (numeric / bool tensors) or a list of bytes objects (DT_STRING tensors)."""
        e = self.entries[key]
        b = self.raw(key, verify)
        if e.dtype == DT_STRING:
            n = int(np.prod(e.shape)) if e.shape else 1
            pos, lens = 0, []
            for _ in range(n):
                ln, pos = _read_varint(b, pos)
                lens.append(ln)
            len_crc = struct.unpack("<I", b[pos:pos + 4])[0]
            pos += 4
            if verify:
                # tensor_bundle.cc WriteStringTensor: the length checksum covers each length as a fixed-width integer (uint32
                # when the string is shorter than 4 GiB), the entry's checksum continues over that masked value and the bytes
                c = 0
                for ln in lens:
                    c = crc32c(struct.pack("<I", ln) if ln <= 0xFFFFFFFF else struct.pack("<Q", ln), c)
                if unmask_crc(len_crc) != c:
                    raise ValueError(f"{key}: string-length crc32c mismatch")
                c = crc32c(struct.pack("<I", len_crc), c)
                c = crc32c(b[pos:], c)
                if unmask_crc(e.crc32c) != c:
                    raise ValueError(f"{key}: string tensor crc32c mismatch")
            out = []
            for ln in lens:
                out.append(bytes(b[pos:pos + ln]))
                pos += ln
            return out
        if e.dtype not in _NP_OF_DT:
            raise ValueError(f"{key}: dtype enum {e.dtype} not supported")
        return np.frombuffer(b, _NP_OF_DT[e.dtype]).reshape(e.shape).copy()


def encode_tensor(val) -> Tuple[int, tuple, bytes, int]:
    """(dtype enum, shape, bytes as stored in a data shard, masked crc32c) of a numpy array or of `bytes` (a scalar DT_STRING
    tensor: varint length, masked crc32c of the length, the bytes -- tensor_bundle.cc WriteStringTensor)."""
    if isinstance(val, (bytes, bytearray)):
        c = crc32c(struct.pack("<I", len(val)))
        len_crc = mask_crc(c)
        c = crc32c(struct.pack("<I", len_crc), c)
        c = crc32c(bytes(val), c)
        return DT_STRING, (), _varint(len(val)) + struct.pack("<I", len_crc) + bytes(val), mask_crc(c)
    a = np.asarray(val)
    if a.dtype not in _DT_OF_NP:
        raise ValueError(f"dtype {a.dtype} not supported")
    blob = np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
    return _DT_OF_NP[a.dtype], a.shape, blob, mask_crc(crc32c(blob))


def write_bundle(prefix: str, tensors: Dict[str, object]):
    """Writes `<prefix>.index` + one data shard.  Values: numpy arrays (float32 / int32 / int64 / bool) or `bytes` (a scalar
    DT_STRING tensor, used for _CHECKPOINTABLE_OBJECT_GRAPH).  Entries are laid out in key order."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data = bytearray()
    items = []
    for key in sorted(tensors):
        try:
            dtype, shape, blob, crc = encode_tensor(tensors[key])
        except ValueError as e:
            raise ValueError(f"{key}: {e}") from None
        items.append((key.encode(), BundleEntry(dtype, shape, 0, len(data), len(blob), crc).serialize()))
        data += blob
    header = _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))          # num_shards = 1, little endian (default), version.producer = 1
    items.append((b"", header))
    with open(_shard_name(prefix, 0, 1), "wb") as f:
        f.write(data)
    write_table(prefix + ".index", items)


# ------------------------------------------------------------------------------------------------ trackable object graph
class ObjNode:
    """One node of TrackableObjectGraph (trackable_object_graph.proto): children (local_name -> node id), attributes
    (name, full_name, checkpoint_key) and optimizer slot references (original variable node, slot name, slot node)."""

    def __init__(self):
        self.children: List[Tuple[str, int]] = []
        self.attributes: List[Tuple[str, str, str]] = []
        self.slots: List[Tuple[int, str, int]] = []


def parse_object_graph(buf: bytes) -> List[ObjNode]:
    nodes = []
    for f, _, v in pb_fields(buf):
        if f != 1:
            continue
        n = ObjNode()
        for f2, _, v2 in pb_fields(v):
            d = pb_fields(v2) if isinstance(v2, bytes) else []
            if f2 == 1:          # ObjectReference { node_id = 1, local_name = 2 }
                g = {ff: vv for ff, _, vv in d}
                n.children.append((g.get(2, b"").decode(), g.get(1, 0)))
            elif f2 == 2:        # SerializedTensor { name = 1, full_name = 2, checkpoint_key = 3 }
                g = {ff: vv for ff, _, vv in d}
                n.attributes.append((g.get(1, b"").decode(), g.get(2, b"").decode(), g.get(3, b"").decode()))
            elif f2 == 3:        # SlotVariableReference { original_variable_node_id = 1, slot_name = 2, slot_variable_node_id = 3 }
                g = {ff: vv for ff, _, vv in d}
                n.slots.append((g.get(1, 0), g.get(2, b"").decode(), g.get(3, 0)))
        nodes.append(n)
    return nodes


def serialize_object_graph(nodes: List[ObjNode]) -> bytes:
    out = b""
    for n in nodes:
        body = b""
        for name, nid in n.children:
            body += _pb_bytes(1, (_pb_varint(1, nid) if nid else b"") + _pb_bytes(2, name.encode()))
        for name, full, key in n.attributes:
            body += _pb_bytes(2, _pb_bytes(1, name.encode()) + (_pb_bytes(2, full.encode()) if full else b"") + _pb_bytes(3, key.encode()))
        for orig, slot, nid in n.slots:
            body += _pb_bytes(3, (_pb_varint(1, orig) if orig else b"") + _pb_bytes(2, slot.encode()) + _pb_varint(3, nid))
        out += _pb_bytes(1, body)
    return out


# ------------------------------------------------------------------------------------------------ the reference's Checkpoint
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_OPT_HYPER = ("iter", "beta_1", "beta_2", "decay", "learning_rate", "momentum_cache")


def keras_layer_order(layer_names: List[str]) -> List[str]:
    """Order in which tf.keras numbers `layer_with_weights-N` for the WDSR functional graph.  Keras sorts a functional model's
    layers by depth (distance from the output, deepest first) and keeps creation order inside one depth.  Counted from the
    output: denorm 0, Add 1, dtsMain / dtsResid 2, reshapeMain / residConv3 3, upscaleConv1 / residConv2 4, last reducer /
    residConv1 5 (modelsTF.py:33-41,45-53,70-73), and the 3-D path is created first (:33 before :34).  So every reducer tail
    (T = 7, 9, 13, 19) gives: mainConv1, the residual blocks, the reducers, residConv1, upscaleConv1, residConv2, residConv3 --
    exactly the order of the shipped checkpoint (checked against its index in tests/test_tfckpt.py)."""
    main = [n for n in layer_names if not n.startswith("residConv")]
    resid = [n for n in layer_names if n.startswith("residConv")]
    out = [n for n in main if n != "upscaleConv1"]
    out += resid[:1] + ["upscaleConv1"] + resid[1:]
    return out


def variable_keys(layer_names: List[str]) -> Dict[str, str]:
    """our variable name `<layer>/{v,g,bias}` -> checkpoint key of tf.train.Checkpoint(model=...) for this graph."""
    keys = {}
    for n, layer in enumerate(keras_layer_order(layer_names)):
        base = f"model/layer_with_weights-{n}"
        keys[f"{layer}/v"] = f"{base}/v{_SUFFIX}"
        keys[f"{layer}/g"] = f"{base}/g{_SUFFIX}"
        keys[f"{layer}/bias"] = f"{base}/layer/bias{_SUFFIX}"
    return keys


def build_object_graph(layer_names: List[str], with_slots: bool = True) -> List[ObjNode]:
    """TrackableObjectGraph of Checkpoint(step, psnr, optimizer, model) (trainClass.py:33-36): root -> {model, optimizer,
    psnr, step, save_counter}; model -> layer_with_weights-N -> {g, v, initialized, layer -> bias}; optimizer -> hyper
    variables + slot references {m, v} for every trainable variable."""
    order = keras_layer_order(layer_names)
    nodes: List[ObjNode] = [ObjNode()]

    def new() -> int:
        nodes.append(ObjNode())
        return len(nodes) - 1

    root = nodes[0]
    model, opt, psnr, step, counter = new(), new(), new(), new(), new()
    root.children = [("model", model), ("optimizer", opt), ("psnr", psnr), ("step", step), ("save_counter", counter)]
    for nid, name in ((psnr, "psnr"), (step, "step"), (counter, "save_counter")):
        nodes[nid].attributes.append(("VARIABLE_VALUE", "Variable" if name != "save_counter" else "save_counter", f"{name}{_SUFFIX}"))
    trainables = []
    for n, layer in enumerate(order):
        ln = new()
        nodes[model].children.append((f"layer_with_weights-{n}", ln))
        base = f"model/layer_with_weights-{n}"
        inner = new()
        nodes[ln].children.append(("layer", inner))
        ids = {}
        for var, full in (("g", f"{layer}/g"), ("v", f"{layer}/kernel"), ("initialized", f"{layer}/initialized")):
            vn = ids[var] = new()
            nodes[ln].children.append((var, vn))
            nodes[vn].attributes.append(("VARIABLE_VALUE", full, f"{base}/{var}{_SUFFIX}"))
            if var != "initialized":
                trainables.append((vn, f"{base}/{var}", full))
        nodes[ln].children.append(("_initialized", ids["initialized"]))      # TFA keeps both names for the same variable
        bn = new()
        nodes[inner].children += [("kernel", ids["v"]), ("bias", bn)]        # the wrapped Conv's kernel IS the wrapper's v
        nodes[bn].attributes.append(("VARIABLE_VALUE", f"{layer}/bias", f"{base}/layer/bias{_SUFFIX}"))
        trainables.append((bn, f"{base}/layer/bias", f"{layer}/bias"))
    for h in _OPT_HYPER:
        hn = new()
        nodes[opt].children.append((h, hn))
        nodes[hn].attributes.append(("VARIABLE_VALUE", f"Nadam/{h}", f"optimizer/{h}{_SUFFIX}"))
    if with_slots:
        for slot in ("m", "v"):
            for vn, base, full in trainables:
                sn = new()
                nodes[opt].slots.append((vn, slot, sn))
                nodes[sn].attributes.append(("VARIABLE_VALUE", f"{full}/{slot}", f"{base}/.OPTIMIZER_SLOT/optimizer/{slot}{_SUFFIX}"))
    return nodes


def save_checkpoint(prefix: str, layer_names: List[str], weights: Dict[str, np.ndarray], step: int, psnr: float, save_counter: int,
                    opt: Optional[dict] = None):
    """Writes one tf.train.Checkpoint of the reference's object graph.  `weights`: `<layer>/{v,g,bias}` -> array (TF layouts,
    modelsTF.py:191-197).  `opt`: {"iter", "learning_rate", "beta_1", "beta_2", "decay", "momentum_cache", "m": {...}, "v": {...}}
    with slot dicts keyed like `weights` (Keras Nadam state, train.py:79-81); None writes no optimizer state."""
    vk = variable_keys(layer_names)
    t: Dict[str, object] = {}
    for name, key in vk.items():
        t[key] = np.asarray(weights[name], np.float32)
    for n in range(len(layer_names)):
        t[f"model/layer_with_weights-{n}/initialized{_SUFFIX}"] = np.asarray(True)        # TFA WeightNormalization._initialized
    t[f"step{_SUFFIX}"] = np.asarray(step, np.int32)
    t[f"psnr{_SUFFIX}"] = np.asarray(psnr, np.float32)
    t[f"save_counter{_SUFFIX}"] = np.asarray(save_counter, np.int64)
    if opt is not None:
        t[f"optimizer/iter{_SUFFIX}"] = np.asarray(opt["iter"], np.int64)
        for h in _OPT_HYPER[1:]:
            t[f"optimizer/{h}{_SUFFIX}"] = np.asarray(opt[h], np.float32)
        for slot in ("m", "v"):
            for name, key in vk.items():
                t[key[:-len(_SUFFIX)] + f"/.OPTIMIZER_SLOT/optimizer/{slot}{_SUFFIX}"] = np.asarray(opt[slot][name], np.float32)
    t[OBJECT_GRAPH_KEY] = serialize_object_graph(build_object_graph(layer_names, with_slots=opt is not None))
    write_bundle(prefix, t)


def load_checkpoint(prefix: str, layer_names: List[str]) -> dict:
    """Reads a tf.train.Checkpoint written by the reference (or by save_checkpoint): {"weights": {...}, "step", "psnr",
    "save_counter", "opt": {...} or None}.  The shipped checkpoints name their reducers convReducer_0..2 in the object graph
    (an older revision of modelsTF.py:160); variables are matched by `layer_with_weights-N` position, so both load."""
    r = BundleReader(prefix)
    vk = variable_keys(layer_names)
    out = {"weights": {name: r.tensor(key) for name, key in vk.items()}}
    for k in ("step", "psnr", "save_counter"):
        out[k] = r.tensor(f"{k}{_SUFFIX}").item() if f"{k}{_SUFFIX}" in r.entries else None
    if f"optimizer/iter{_SUFFIX}" in r.entries:
        o = {"iter": int(r.tensor(f"optimizer/iter{_SUFFIX}"))}
        for h in _OPT_HYPER[1:]:
            if f"optimizer/{h}{_SUFFIX}" in r.entries:
                o[h] = float(r.tensor(f"optimizer/{h}{_SUFFIX}"))
        for slot in ("m", "v"):
            o[slot] = {}
            for name, key in vk.items():
                sk = key[:-len(_SUFFIX)] + f"/.OPTIMIZER_SLOT/optimizer/{slot}{_SUFFIX}"
                if sk in r.entries:
                    o[slot][name] = r.tensor(sk)
        out["opt"] = o
    else:
        out["opt"] = None
    return out


# ------------------------------------------------------------------------------------------------ CheckpointManager state file
def read_checkpoint_state(ckpt_dir: str) -> dict:
    """Parses the text-format CheckpointState file `checkpoint` (trainClass.py:37-39 CheckpointManager)."""
    st = {"model_checkpoint_path": None, "all_model_checkpoint_paths": [], "all_model_checkpoint_timestamps": [], "last_preserved_timestamp": None}
    p = os.path.join(ckpt_dir, "checkpoint")
    if not os.path.exists(p):
        return st
    for line in open(p):
        if ":" not in line:
            continue
        k, v = line.split(":", 1)
        k, v = k.strip(), v.strip()
        if v.startswith('"'):
            v = v[1:-1]
        if k == "model_checkpoint_path":
            st[k] = v
        elif k == "all_model_checkpoint_paths":
            st[k].append(v)
        elif k == "all_model_checkpoint_timestamps":
            st[k].append(float(v))
        elif k == "last_preserved_timestamp":
            st[k] = float(v)
    return st


def write_checkpoint_state(ckpt_dir: str, paths: List[str], timestamps: List[float], last_preserved: Optional[float] = None):
    lines = [f'model_checkpoint_path: "{paths[-1]}"'] + [f'all_model_checkpoint_paths: "{p}"' for p in paths]
    lines += [f"all_model_checkpoint_timestamps: {t!r}" for t in timestamps]
    lines.append(f"last_preserved_timestamp: {(last_preserved if last_preserved is not None else time.time())!r}")
    with open(os.path.join(ckpt_dir, "checkpoint"), "w") as f:
        f.write("\n".join(lines) + "\n")


def latest_checkpoint(ckpt_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint / CheckpointManager.latest_checkpoint: prefix of the newest checkpoint, or None."""
    st = read_checkpoint_state(ckpt_dir)
    name = st["model_checkpoint_path"]
    if not name:
        return None
    prefix = name if os.path.isabs(name) else os.path.join(ckpt_dir, name)
    return prefix if os.path.exists(prefix + ".index") else None
