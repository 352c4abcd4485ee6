"""TensorBoard event files for the training log, in the format `tf.summary.create_file_writer` / `tf.summary.scalar` write.

The reference logs four scalars through a TF2 summary writer on `logDir` (models/trainClass.py:41,99-104,112-116): tags
'Train PSNR' and 'Train loss' every step, 'Test loss' and 'Test PSNR' at every evaluation, all at step = the global step.
An event file `events.out.tfevents.<time>.<host>.<pid>.<uid>.v2` is a TFRecord stream (uint64 length, masked crc32c of the
length, payload, masked crc32c of the payload) of Event protos: first {wall_time, file_version "brain.Event:2"}, then per
scalar {wall_time, step, summary{value{tag, tensor{DT_FLOAT, scalar shape, 4 content bytes}, metadata{plugin "scalars"}}}}.

Pinned by the reference's own log files: tests/test_tbevents.py re-encodes every record of two event files shipped under
modelInfo/logs_p16t9c85r12/NIR/ (fixtures in tests/golden/tb_events/) and requires the bytes to match, and reads a file
written here back with the `tensorboard` package's own loader.  Host-side only (pure Python)."""
from __future__ import annotations

import os
import socket
import struct
import time
from typing import Iterator, List, NamedTuple, Optional

from .tfckpt import _pb_bytes, _pb_varint, _varint, crc32c, mask_crc, pb_fields, unmask_crc


class ScalarEvent(NamedTuple):
    wall_time: float
    step: int
    tag: Optional[str]           # None for the file-version record
    value: Optional[float]
    file_version: Optional[str] = None


def _pb_double(f: int, v: float) -> bytes:
    return _varint((f << 3) | 1) + struct.pack("<d", v)


def encode_event(ev: ScalarEvent) -> bytes:
    out = _pb_double(1, ev.wall_time)
    if ev.file_version is not None:
        return out + _pb_bytes(3, ev.file_version.encode())
    if ev.step:
        out += _pb_varint(2, ev.step)
    tensor = _pb_varint(1, 1) + _pb_bytes(2, b"") + _pb_bytes(4, struct.pack("<f", ev.value))       # DT_FLOAT, rank 0, content
    meta = _pb_bytes(1, _pb_bytes(1, b"scalars"))                                                     # SummaryMetadata.plugin_data.plugin_name
    value = _pb_bytes(1, ev.tag.encode()) + _pb_bytes(8, tensor) + _pb_bytes(9, meta)
    return out + _pb_bytes(5, _pb_bytes(1, value))


def decode_event(buf: bytes) -> ScalarEvent:
    wall, step, tag, val, ver = 0.0, 0, None, None, None
    for f, _, v in pb_fields(buf):
        if f == 1:
            wall = struct.unpack("<d", v)[0]
        elif f == 2:
            step = v
        elif f == 3:
            ver = v.decode()
        elif f == 5:
            for f2, _, v2 in pb_fields(v):
                if f2 != 1:
                    continue
                for f3, _, v3 in pb_fields(v2):
                    if f3 == 1:
                        tag = v3.decode()
                    elif f3 == 2:                      # TF1-style simple_value
                        val = struct.unpack("<f", v3)[0]
                    elif f3 == 8:
                        t = {ff: vv for ff, _, vv in pb_fields(v3)}
                        if 4 in t and len(t[4]) == 4:
                            val = struct.unpack("<f", t[4])[0]
                        elif 5 in t:
                            val = struct.unpack("<f", t[5][:4])[0]
    return ScalarEvent(wall, step, tag, val, ver)


def frame_record(payload: bytes) -> bytes:
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", mask_crc(crc32c(head))) + payload + struct.pack("<I", mask_crc(crc32c(payload)))


def read_records(path: str, verify: bool = True) -> Iterator[bytes]:
    b = open(path, "rb").read()
    pos = 0
    while pos + 12 <= len(b):
        (ln,) = struct.unpack("<Q", b[pos:pos + 8])
        if ln and pos + 16 + ln > len(b):
            break                                      # truncated tail (a writer that was killed): what TF's reader tolerates too
        payload = b[pos + 12:pos + 12 + ln]
        if ln == 0 and not any(b[pos:]):
            break                                      # zero-filled tail: blocks the file system pre-allocated for a writer that
                                                       # died (one of the reference's own logs ends like this)
        if verify:
            if unmask_crc(struct.unpack("<I", b[pos + 8:pos + 12])[0]) != crc32c(b[pos:pos + 8]):
                raise ValueError(f"{path}: length crc mismatch at {pos}")
            if unmask_crc(struct.unpack("<I", b[pos + 12 + ln:pos + 16 + ln])[0]) != crc32c(payload):
                raise ValueError(f"{path}: payload crc mismatch at {pos}")
        yield payload
        pos += 16 + ln


def read_scalars(path: str, verify: bool = True) -> List[ScalarEvent]:
    return [decode_event(r) for r in read_records(path, verify)]


class SummaryWriter:
    """tf.summary.create_file_writer(logdir) + writer.as_default() scalars (trainClass.py:41,99-116)."""

    _uid = 0

    def __init__(self, logdir: str, filename_suffix: str = ".v2"):
        os.makedirs(logdir, exist_ok=True)
        SummaryWriter._uid += 1
        now = time.time()
        name = f"events.out.tfevents.{int(now)}.{socket.gethostname()}.{os.getpid()}.{SummaryWriter._uid}{filename_suffix}"
        self.path = os.path.join(logdir, name)
        self._f = open(self.path, "ab")
        self._f.write(frame_record(encode_event(ScalarEvent(now, 0, None, None, "brain.Event:2"))))
        self._f.flush()

    def scalar(self, tag: str, value: float, step: int, wall_time: float = None):
        ev = ScalarEvent(time.time() if wall_time is None else wall_time, int(step), tag, float(value))
        self._f.write(frame_record(encode_event(ev)))

    def flush(self):
        self._f.flush()

    def close(self):
        if self._f:
            self._f.close()
            self._f = None
