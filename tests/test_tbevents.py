"""TensorBoard event-file compatibility of the training log, pinned by the reference's OWN log files.

tests/golden/tb_events/ holds two event files copied verbatim from the reference repository
(modelInfo/logs_p16t9c85r12/NIR/events.out.tfevents.1583765479... and ...1583765851...) and the first 2010 records of
...1583836514... (cut at a record boundary), all written by tf.summary in models/trainClass.py:41,99-116."""
import glob
import math
import os

import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tb_events")
TAGS = {"Train PSNR", "Train loss", "Test loss", "Test PSNR"}          # trainClass.py:101-102,113-114


@pytest.fixture(scope="module")
def tb():
    from probav_b200 import tbevents
    return tbevents


def test_reference_logs_decode_and_reencode_byte_exact(tb):
    files = sorted(glob.glob(os.path.join(GOLD, "events.out.tfevents.*")))
    assert len(files) == 3
    seen = set()
    for p in files:
        recs = list(tb.read_records(p))            # verifies both crc32c fields of every record
        evs = [tb.decode_event(r) for r in recs]
        assert evs[0].file_version == "brain.Event:2"
        for e, r in zip(evs, recs):
            assert tb.encode_event(e) == r
        assert b"".join(tb.frame_record(tb.encode_event(e)) for e in evs) == open(p, "rb").read()
        seen |= {e.tag for e in evs[1:]}
    assert seen == TAGS
    # the converged run: cadence and value ranges quoted in SURVEY.md section 6
    evs = tb.read_scalars(files[-1])
    train = [e for e in evs if e.tag == "Train PSNR"]
    assert [e.step for e in train[:3]] == [163018, 163019, 163020]          # one Train record per global step
    test = [e for e in evs if e.tag == "Test PSNR"]
    assert test and all(40.0 < e.value < 55.0 for e in test)
    tl = [e for e in evs if e.tag == "Test loss"]
    assert [e.step for e in tl] == [e.step for e in test]                   # evaluation writes both at the same step


def test_writer_output_is_read_by_tensorboards_own_loader(tb, tmp_path):
    w = tb.SummaryWriter(str(tmp_path))
    vals = [(1, "Train PSNR", 41.25), (1, "Train loss", 612.5), (2, "Train PSNR", 42.0), (2, "Train loss", 600.0),
            (2, "Test loss", 590.0), (2, "Test PSNR", float("nan"))]
    for step, tag, v in vals:
        w.scalar(tag, v, step)
    w.close()
    assert os.path.basename(w.path).startswith("events.out.tfevents.") and w.path.endswith(".v2")
    back = tb.read_scalars(w.path)
    assert back[0].file_version == "brain.Event:2"
    assert [(e.step, e.tag) for e in back[1:]] == [(s, t) for s, t, _ in vals]
    loader = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")
    from tensorboard.util import tensor_util
    got = []
    for ev in loader.EventFileLoader(w.path).Load():
        for v in ev.summary.value:
            assert v.metadata.plugin_data.plugin_name == "scalars"
            got.append((ev.step, v.tag, float(tensor_util.make_ndarray(v.tensor))))
    assert len(got) == len(vals)
    for (s, t, v), (s2, t2, v2) in zip(vals, got):
        assert (s, t) == (s2, t2) and (v == v2 or (math.isnan(v) and math.isnan(v2)))


def test_truncated_tail_is_tolerated_and_corruption_is_caught(tb, tmp_path):
    src = sorted(glob.glob(os.path.join(GOLD, "events.out.tfevents.*")))[0]
    b = open(src, "rb").read()
    p = tmp_path / "cut.v2"
    p.write_bytes(b[:-7])                                  # a writer killed mid-record
    assert len(list(tb.read_records(str(p)))) == len(list(tb.read_records(src))) - 1
    bad = bytearray(b)
    bad[40] ^= 1
    p.write_bytes(bad)
    with pytest.raises(ValueError, match="crc"):
        list(tb.read_records(str(p)))


def test_event_codec_agrees_with_tensorflows_generated_proto(tb):
    """encode_event against TensorFlow's generated Event / Summary classes (shipped inside `tensorboard`): parse, compare
    fields, and re-serialize byte for byte."""
    pb2 = pytest.importorskip("tensorboard.compat.proto.event_pb2")
    for ev in (tb.ScalarEvent(1583836565.721661, 163018, "Train PSNR", 47.63618087768555),
               tb.ScalarEvent(12.5, 0, "Test loss", float("inf")),
               tb.ScalarEvent(1583765479.0, 0, None, None, "brain.Event:2")):
        blob = tb.encode_event(ev)
        msg = pb2.Event.FromString(blob)
        assert msg.SerializeToString() == blob
        assert msg.wall_time == ev.wall_time and msg.step == ev.step
        if ev.tag is None:
            assert msg.file_version == "brain.Event:2"
        else:
            v = msg.summary.value[0]
            assert v.tag == ev.tag and v.metadata.plugin_data.plugin_name == "scalars"
            assert v.tensor.dtype == 1 and len(v.tensor.tensor_shape.dim) == 0            # DT_FLOAT scalar
            import struct
            assert struct.unpack("<f", v.tensor.tensor_content)[0] == pytest.approx(ev.value, rel=1e-7) or math.isinf(ev.value)
