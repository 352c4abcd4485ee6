"""TF checkpoint (tensor bundle) compatibility, pinned by the reference's OWN checkpoint files.

tests/golden/tf_ckpt/ holds three files copied verbatim from the reference repository
(modelInfo/ckpt_p16t9c85r12/NIR/{ckpt-124.index, ckpt-124.data-00000-of-00002, checkpoint}; shard 1 with the weights is
not shipped there, .MISSING_LARGE_BLOBS).  They were written by TensorFlow's tf.train.CheckpointManager in
models/trainClass.py:33-39,118-120, so they are golden vectors for the on-disk format."""
import os

import numpy as np
import pytest

from oracle.wdsr import init_params, layer_specs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tf_ckpt")
PREFIX = os.path.join(GOLD, "ckpt-124")


@pytest.fixture(scope="module")
def tfckpt():
    from probav_b200 import tfckpt as t
    return t


@pytest.fixture(scope="module")
def specs():
    return layer_specs(3, 32, (3, 3, 3), 12, 8, 0.8, 9, True)       # cfg/p16t9c85r12.cfg [Net]


def test_crc32c_known_answers(tfckpt):
    assert tfckpt.crc32c(b"123456789") == 0xE3069283               # the CRC-32C check value (RFC 3720 appendix B.4)
    assert tfckpt.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tfckpt.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfckpt.crc32c(b"6789", tfckpt.crc32c(b"12345")) == 0xE3069283
    for c in (0, 1, 0xDEADBEEF, 0xFFFFFFFF):
        assert tfckpt.unmask_crc(tfckpt.mask_crc(c)) == c
    # ... and against the TFRecord checksum implementation that ships with the `tensorboard` package
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 64, 1000, 4097):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert pw.masked_crc32c(b) == tfckpt.mask_crc(tfckpt.crc32c(b)), n


def test_reference_index_parses_and_every_checksum_holds(tfckpt, specs):
    r = tfckpt.BundleReader(PREFIX)                    # verifies the crc32c of every table block
    assert (r.num_shards, r.endianness, r.version) == (2, 0, (1, 0))
    assert len(r.entries) == 450                       # 44 layers x (v, g, bias) x (value, m, v) + 44 initialized + 6 + 3 + graph
    # shard 0 is shipped: its tensors must decode and match their stored checksums (numeric and string paths)
    assert int(r.tensor("step/.ATTRIBUTES/VARIABLE_VALUE")) == 285107
    assert int(r.tensor("save_counter/.ATTRIBUTES/VARIABLE_VALUE")) == 124
    graph = r.tensor(tfckpt.OBJECT_GRAPH_KEY)
    assert len(graph) == 1 and len(graph[0]) == 64048
    with pytest.raises(FileNotFoundError):
        r.tensor("psnr/.ATTRIBUTES/VARIABLE_VALUE")    # lives in the shard the reference does not ship
    # every variable of the p16t9c85r12 graph is where variable_keys() says, with the layer's shape
    vk = tfckpt.variable_keys([s["name"] for s in specs])
    for s in specs:
        e = r.entries[vk[s["name"] + "/v"]]
        assert e.dtype == tfckpt.DT_FLOAT and e.shape == (*s["k"], s["cin"], s["cout"]), s["name"]
        assert r.entries[vk[s["name"] + "/g"]].shape == (s["cout"],)
        assert r.entries[vk[s["name"] + "/bias"]].shape == (s["cout"],)
    assert sum(int(np.prod(r.entries[k].shape)) for k in vk.values()) == 535267


def test_writers_are_byte_exact_on_the_reference_files(tfckpt, tmp_path):
    """Re-encoding what was parsed must give back TensorFlow's bytes: every BundleEntryProto, the whole index table
    (blocks, restart points, shortened index key, footer) and the data shard (numeric + string tensor encodings)."""
    items = tfckpt.read_table(PREFIX + ".index")
    for k, v in items:
        if k:
            assert tfckpt.BundleEntry.parse(v).serialize() == v, k
    out = tmp_path / "again.index"
    tfckpt.write_table(str(out), items)
    assert out.read_bytes() == open(PREFIX + ".index", "rb").read()
    r = tfckpt.BundleReader(PREFIX)
    shard = open(PREFIX + ".data-00000-of-00002", "rb").read()
    vals = {"step/.ATTRIBUTES/VARIABLE_VALUE": np.asarray(285107, np.int32),
            "save_counter/.ATTRIBUTES/VARIABLE_VALUE": np.asarray(124, np.int64),
            tfckpt.OBJECT_GRAPH_KEY: r.tensor(tfckpt.OBJECT_GRAPH_KEY)[0]}
    rebuilt = bytearray(len(shard))
    for k, val in vals.items():
        e = r.entries[k]
        dtype, shape, blob, crc = tfckpt.encode_tensor(val)
        assert (dtype, tuple(shape), len(blob), crc) == (e.dtype, e.shape, e.size, e.crc32c), k
        rebuilt[e.offset:e.offset + e.size] = blob
    assert bytes(rebuilt) == shard
    # and the object graph proto round-trips through the node model
    g = vals[tfckpt.OBJECT_GRAPH_KEY]
    assert tfckpt.serialize_object_graph(tfckpt.parse_object_graph(g)) == g


def _paths(tfckpt, nodes):
    """checkpoint_key -> set of local-name paths from the root, plus (variable key, slot name) -> slot key."""
    paths, slots = {}, {}
    key_of = {i: n.attributes[0][2] for i, n in enumerate(nodes) if n.attributes}
    stack, seen = [(0, ())], set()
    while stack:
        nid, path = stack.pop()
        if (nid, path) in seen or len(path) > 6:
            continue
        seen.add((nid, path))
        if nid in key_of:
            paths.setdefault(key_of[nid], set()).add("/".join(path))
        for name, child in nodes[nid].children:
            stack.append((child, path + (name,)))
    for n in nodes:
        for orig, slot, sn in n.slots:
            slots[(key_of[orig], slot)] = key_of[sn]
    return paths, slots


def test_generated_object_graph_matches_the_reference_topology(tfckpt, specs):
    ref_nodes = tfckpt.parse_object_graph(tfckpt.BundleReader(PREFIX).tensor(tfckpt.OBJECT_GRAPH_KEY)[0])
    ours = tfckpt.build_object_graph([s["name"] for s in specs])
    rp, rs = _paths(tfckpt, ref_nodes)
    op, os_ = _paths(tfckpt, ours)
    assert set(op) == set(rp)                              # the same checkpoint keys hang off the graph
    for key, p in op.items():
        assert p <= rp[key], (key, p - rp[key])            # ... under names TensorFlow's restore walks too
    assert os_ == rs                                       # the same 264 optimizer slot references
    # variable full names agree except for the reducers the shipped checkpoint still calls convReducer_0..2
    rf = {n.attributes[0][2]: n.attributes[0][1] for n in ref_nodes if n.attributes}
    of = {n.attributes[0][2]: n.attributes[0][1] for n in ours if n.attributes}
    diff = {k for k in of if of[k] != rf[k]}
    assert all("convReducer_" in of[k] for k in diff) and len(diff) == 3 * 4 * 3 - 3 * 2


def test_saved_bundle_has_the_reference_key_set_and_round_trips(tfckpt, specs, tmp_path):
    names = [s["name"] for s in specs]
    w = {k: v.numpy().astype(np.float32) for k, v in init_params(specs, seed=3).items()}
    rng = np.random.default_rng(0)
    opt = {"iter": 285107, "learning_rate": 5e-4, "beta_1": 0.9, "beta_2": 0.999, "decay": 0.0, "momentum_cache": 0.25,
           "m": {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in w.items()},
           "v": {k: rng.random(v.shape).astype(np.float32) for k, v in w.items()}}
    prefix = str(tmp_path / "ckpt-1")
    tfckpt.save_checkpoint(prefix, names, w, step=285107, psnr=47.5, save_counter=1, opt=opt)
    ref, got = tfckpt.BundleReader(PREFIX), tfckpt.BundleReader(prefix)
    assert got.keys() == ref.keys()
    for k in ref.keys():
        if k != tfckpt.OBJECT_GRAPH_KEY:
            assert (got.entries[k].dtype, got.entries[k].shape, got.entries[k].size) == \
                   (ref.entries[k].dtype, ref.entries[k].shape, ref.entries[k].size), k
    back = tfckpt.load_checkpoint(prefix, names)
    assert back["step"] == 285107 and back["save_counter"] == 1 and abs(back["psnr"] - 47.5) < 1e-6
    for k in w:
        assert np.array_equal(back["weights"][k], w[k]), k
        assert np.array_equal(back["opt"]["m"][k], opt["m"][k]) and np.array_equal(back["opt"]["v"][k], opt["v"][k])
    assert back["opt"]["iter"] == 285107 and abs(back["opt"]["momentum_cache"] - 0.25) < 1e-7
    # a flipped data byte is caught by the tensor checksum
    shard = prefix + ".data-00000-of-00001"
    b = bytearray(open(shard, "rb").read())
    b[got.entries["model/layer_with_weights-5/v/.ATTRIBUTES/VARIABLE_VALUE"].offset + 7] ^= 0x40
    open(shard, "wb").write(b)
    with pytest.raises(ValueError, match="crc32c"):
        tfckpt.BundleReader(prefix).tensor("model/layer_with_weights-5/v/.ATTRIBUTES/VARIABLE_VALUE")


def test_multi_block_tables_and_other_reducer_tails(tfckpt, tmp_path, monkeypatch):
    monkeypatch.setattr(tfckpt, "_BLOCK_SIZE", 700)
    items = tfckpt.read_table(PREFIX + ".index")
    p = str(tmp_path / "small_blocks.index")
    tfckpt.write_table(p, items)
    assert tfckpt.read_table(p) == items
    for T, nred in ((7, 2), (13, 5)):
        names = [s["name"] for s in layer_specs(3, 32, (3, 3, 3), 2, 8, 0.8, T, True)]
        order = tfckpt.keras_layer_order(names)
        assert order[-4:] == ["residConv1", "upscaleConv1", "residConv2", "residConv3"]
        assert order[-4 - nred:-4] == [f"convReducer_{i + 1}" for i in range(nred)]


def test_checkpoint_state_file(tfckpt, tmp_path):
    st = tfckpt.read_checkpoint_state(GOLD)
    assert st["model_checkpoint_path"] == "ckpt-124"
    assert st["all_model_checkpoint_paths"] == [f"ckpt-{i}" for i in range(120, 125)]
    assert len(st["all_model_checkpoint_timestamps"]) == 5
    assert tfckpt.latest_checkpoint(GOLD) == os.path.join(GOLD, "ckpt-124")
    tfckpt.write_checkpoint_state(str(tmp_path), st["all_model_checkpoint_paths"], st["all_model_checkpoint_timestamps"],
                                  st["last_preserved_timestamp"])
    assert open(tmp_path / "checkpoint").read() == open(os.path.join(GOLD, "checkpoint")).read()


REF_ROOT = "/root/reference/modelInfo"


@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason="the reference checkout is only present in the build container")
def test_every_reference_checkpoint_and_log_reencodes_byte_exact(tfckpt, tmp_path):
    """All 11 checkpoint indices (NIR 40-43, 120-124; RED 124, 128), their shard-0 files and all event files the reference
    repository ships: every checksum holds and the writers reproduce the files byte for byte."""
    import glob
    from probav_b200 import tbevents
    idx = sorted(glob.glob(os.path.join(REF_ROOT, "ckpt_*", "*", "*.index")))
    assert len(idx) >= 11
    for p in idx:
        prefix = p[:-len(".index")]
        items = tfckpt.read_table(p)
        out = str(tmp_path / "again.index")
        tfckpt.write_table(out, items)
        assert open(out, "rb").read() == open(p, "rb").read(), p
        r = tfckpt.BundleReader(prefix)
        assert len(r.entries) == 450
        if r.has_shard(0):
            step = int(r.tensor("step/.ATTRIBUTES/VARIABLE_VALUE"))
            counter = int(r.tensor("save_counter/.ATTRIBUTES/VARIABLE_VALUE"))
            assert counter == int(prefix.rsplit("-", 1)[1]) and step > 0
            g = r.tensor(tfckpt.OBJECT_GRAPH_KEY)[0]
            assert tfckpt.serialize_object_graph(tfckpt.parse_object_graph(g)) == g
    logs = sorted(glob.glob(os.path.join(REF_ROOT, "logs_*", "*", "events.out.tfevents.*")))
    assert len(logs) >= 9
    n = 0
    for p in logs:
        raw = open(p, "rb").read()
        recs = list(tbevents.read_records(p))
        again = b"".join(tbevents.frame_record(tbevents.encode_event(tbevents.decode_event(r))) for r in recs)
        assert raw[:len(again)] == again and not any(raw[len(again):]), p      # one log ends in a zero-filled tail (dead writer)
        n += len(recs)
    assert n > 200000          # ~14 MB of scalars: the converged NIR and RED runs


def test_object_graph_codec_agrees_with_tensorflows_generated_proto(tfckpt, specs):
    """The hand-written protobuf codec against TensorFlow's own generated TrackableObjectGraph class (shipped inside the
    `tensorboard` package): same nodes from the reference's graph, and a graph generated here parses and re-serializes
    byte for byte through it."""
    pb2 = pytest.importorskip("tensorboard.compat.proto.trackable_object_graph_pb2")
    ref_bytes = tfckpt.BundleReader(PREFIX).tensor(tfckpt.OBJECT_GRAPH_KEY)[0]
    ours_bytes = tfckpt.serialize_object_graph(tfckpt.build_object_graph([s["name"] for s in specs]))
    for blob in (ref_bytes, ours_bytes):
        msg = pb2.TrackableObjectGraph.FromString(blob)
        mine = tfckpt.parse_object_graph(blob)
        assert len(msg.nodes) == len(mine)
        for a, b in zip(msg.nodes, mine):
            assert [(c.local_name, c.node_id) for c in a.children] == b.children
            assert [(t.name, t.full_name, t.checkpoint_key) for t in a.attributes] == b.attributes
            assert [(s.original_variable_node_id, s.slot_name, s.slot_variable_node_id) for s in a.slot_variables] == b.slots
        assert msg.SerializeToString() == blob
    assert len(pb2.TrackableObjectGraph.FromString(ours_bytes).nodes) == 1 + 5 + 44 * 6 + 6 + 264
