"""CPU: host-side logic -- cfg grammar, input pipeline, data-parallel plumbing (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import probav_b200 as pb
from probav_b200 import parallel
from probav_b200.trainClass import Mean, batched, shuffled_index_stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parse_config_p16t9c85r12():
    c = pb.parseConfig(os.path.join(ROOT, "cfg", "p16t9c85r12.cfg"))
    assert (c["num_res_blocks"], c["num_low_res_imgs"], c["scale"], c["num_filters"], c["kernel_size"], c["exp_rate"]) == (12, 9, 3, 32, 3, 8)
    assert c["decay_rate"] == 0.8 and c["is_grayscale"] is True
    assert c["max_shift"] == 6 and c["patch_size"] == 16
    assert c["optimizer"] == "nadam" and c["loss"] == "l1" and c["batch_size"] == 128 and c["learning_rate"] == 0.0005
    assert isinstance(c["ckpt"], list) and isinstance(c["raw_data"], str)
    # '.cfg' appended and cfg/ searched (parseConfig.py:10-13)
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        assert pb.parseConfig("p16t12c85r12")["num_low_res_imgs"] == 9
    finally:
        os.chdir(cwd)


def test_parse_config_rejects_unknown_field(tmp_path):
    f = tmp_path / "bad.cfg"
    f.write_text("[Directories]\nraw_data=x\n[Net]\nnum_filters=32\nbogus_field=3\n")
    with pytest.raises(AssertionError):
        pb.parseConfig(str(f))
    g = tmp_path / "ok.cfg"
    g.write_text("# comment\n[Directories]\nanything_goes=here\n\n[Train]\nlearning_rate=0.1\nloss=l2\n[Preprocessing]\nto_flip=1\nlow_res_patch_thresholds=0.5,0.6\n")
    c = pb.parseConfig(str(g))
    assert c["anything_goes"] == "here" and c["learning_rate"] == 0.1 and c["to_flip"] is True and c["low_res_patch_thresholds"] == [0.5, 0.6]


def test_shuffle_stream_is_a_permutation_per_epoch_and_last_batch_is_partial():
    rng = np.random.default_rng(0)
    idx = list(shuffled_index_stream(100, 2, 16, rng))
    assert sorted(idx[:100]) == list(range(100)) and sorted(idx[100:]) == list(range(100))
    # a shuffle buffer of 16 cannot emit element k before position k-15
    assert all(v <= pos + 15 for pos, v in enumerate(idx[:100]))
    b = list(batched(iter(range(10)), 4))
    assert [len(x) for x in b] == [4, 4, 2]          # no drop_remainder (utils.py:32-34)


def test_mean_metric_and_shard_bounds():
    m = Mean()
    m(2.0); m(4.0)
    assert m.result() == 3.0
    m.reset_states()
    assert m.result() == 0.0
    for n in (0, 1, 7, 128, 129):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_loss_selector_and_optimizer_selector():
    L = pb.Losses((48, 48, 1))
    assert pb.loss_from_config(L, "l1").__name__ == "shiftCompensatedL1Loss"
    assert pb.loss_from_config(L, "sobel_l1_mix").__name__ == "shiftCompensatedL1EdgeLoss"
    with pytest.raises(ValueError):
        pb.loss_from_config(L, "nope")
    assert pb.optimizers.from_config("nadam", 5e-4).kind == "nadam"
    assert pb.optimizers.from_config("adam", 5e-4).kind == "adam"
    assert pb.optimizers.from_config("rmsprop", 5e-4).kind == "sgd"      # anything else -> SGD (train.py:82-83)
    assert L.cropSizeHeight == 42 and L.maxPixelShift == 6


# ------------------------------------------------------------------------------------------- gloo, world_size 2
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, ws, port, q, nsamp=5):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, ws) and parallel.world() == (rank, ws)
    # the DP contract on the oracle: per-rank grads of sum_b loss_b / global_batch, SUM all-reduced == full-batch grads
    from oracle.losses import OracleLosses
    from oracle.step import loss_and_grads
    from oracle.wdsr import OracleWDSR, init_params
    from probav_b200 import synth
    om = OracleWDSR(8075.2045, 3160.7272, 6, 3, 8, (3, 3, 3), 1, 2, 0.8, 9, 16)
    p = init_params(om.specs, seed=0)
    lr, hr, mask = synth.make_batch(nsamp, seed=3, hr_zero_under_mask=True)      # 5 samples -> unequal shards 3 + 2
    ol = OracleLosses((48, 48, 1))
    lo, hi = parallel.shard_bounds(nsamp, rank, ws)
    t = lambda a: torch.from_numpy(a[lo:hi])
    n_local = hi - lo
    if n_local == 0:
        # a partial last global batch with fewer samples than ranks: the empty rank contributes zero gradients and zero-weight
        # metrics but joins every collective (trainClass._dp_step with B = 0; pv_train_forward_backward_staged zero-fills)
        flat = torch.zeros(sum(v.numel() for v in p.values()), dtype=torch.float64)
        loss, cps = torch.zeros(()), torch.zeros(1)
    else:
        loss, g, _, cps = loss_and_grads(om, ol, p, t(lr).double(), t(hr).double(), t(mask))
        flat = torch.cat([v.reshape(-1) for v in g.values()]) * (n_local * parallel.grad_scale(nsamp))   # mean over shard -> sum/global
    parallel.allreduce_sum_(flat)
    gl, gc = parallel.reduce_metrics(float(loss), float(cps.mean()), n_local)
    w0 = torch.zeros(4) + rank
    parallel.broadcast_(w0, 0)
    if rank == 0:
        full_loss, gfull, _, cfull = loss_and_grads(om, ol, p, torch.from_numpy(lr).double(), torch.from_numpy(hr).double(), torch.from_numpy(mask))
        ref = torch.cat([v.reshape(-1) for v in gfull.values()])
        q.put((float((flat - ref).abs().max() / ref.abs().max()), abs(gl - float(full_loss)), abs(gc - float(cfull.mean())), float(w0.sum())))
    else:
        q.put(("r1", float(w0.sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nsamp", [5, 1])       # 1 sample on 2 ranks: rank 1's shard is empty (last partial batch < world size)
def test_data_parallel_contract_gloo_world2(nsamp):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q, nsamp)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    main = [r for r in res if r[0] != "r1"][0]
    other = [r for r in res if r[0] == "r1"][0]
    assert main[0] < 1e-10 and main[1] < 1e-9 and main[2] < 1e-9
    assert main[3] == 0.0 and other[1] == 0.0           # broadcast from rank 0


def test_prefetch_loader_yields_the_rank_shard_of_every_global_batch():
    """pipeline.PrefetchLoader on host arrays (device=None): same batches as gathering by hand, for both ranks of a
    2-rank job, including the partial last batch (utils/utils.py:32-34 has no drop_remainder)."""
    from probav_b200 import parallel
    from probav_b200.pipeline import PrefetchLoader
    from probav_b200.trainClass import batched, shuffled_index_stream
    N = 50
    X = np.arange(N * 6, dtype=np.float32).reshape(N, 2, 3)
    y = np.arange(N, dtype=np.float32).reshape(N, 1)
    m = (np.arange(N) % 3 == 0).reshape(N, 1)
    ref = list(batched(shuffled_index_stream(N, 2, 16, np.random.default_rng(0)), 8))
    assert sum(len(b) for b in ref) == 2 * N and len(ref[-1]) == 4
    for rank in (0, 1):
        ld = PrefetchLoader((X, y, m), batched(shuffled_index_stream(N, 2, 16, np.random.default_rng(0)), 8), device=None,
                            rank=rank, world_size=2, depth=2)
        seen = 0
        for (gb, (a, b, c)), idx in zip(ld, ref):
            lo, hi = parallel.shard_bounds(len(idx), rank, 2)
            sel = np.sort(idx[lo:hi])
            assert gb == len(idx)
            assert np.array_equal(a, X[sel]) and np.array_equal(b, y[sel]) and np.array_equal(c, m[sel].view(np.uint8))
            seen += 1
        assert seen == len(ref)
    # leaving the loop early stops the producer thread
    ld = PrefetchLoader((X, y, m), batched(shuffled_index_stream(N, 50, 16, np.random.default_rng(1)), 8), device=None)
    for i, _ in enumerate(ld):
        if i == 2:
            break
    ld.close()
    assert ld._thread is None


def test_prefetch_loader_propagates_producer_errors():
    from probav_b200.pipeline import PrefetchLoader
    X = np.zeros((4, 2), np.float32)
    ld = PrefetchLoader((X,), [np.array([0, 1]), np.array([7, 9])], device=None)
    with pytest.raises(IndexError):
        for _ in ld:
            pass


def test_cli_flags_and_png_writer(tmp_path):
    """train.py / test.py keep the reference's flags and defaults (train.py:26-32, test.py:25-31); predictions are saved as
    16-bit grayscale PNGs (test.py:99)."""
    from probav_b200 import cli
    a = cli.train_parser().parse_args([])
    assert (a.cfg, a.band, a.modelType) == ("cfg/yourcfg.cfg", "NIR", "patchNet")
    b = cli.test_parser().parse_args(["--totest", "TRAIN"])
    assert (b.cfg, b.band, b.totest) == ("cfg/FINAL.cfg", "RED", "TRAIN")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 65536, size=(37, 53)).astype(np.float64)
    p = str(tmp_path / "imgset0001.png")
    cli.write_png16(p, img)
    back = cli.read_png16(p)
    assert back.dtype == np.uint16 and np.array_equal(back, img.astype(np.uint16))
    with pytest.raises(ValueError):
        cli.write_png16(p, np.zeros((2, 2, 3)))


def test_predict_helpers_with_a_stub_model():
    """Host-side logic of the predict path (test.py:125-160) against a stub model: batch splitting with a remainder, the
    n x n row-major stitch, and the frame-order test-time augmentation (a frame-order-invariant model must be unchanged)."""
    import probav_b200 as pb

    class Stub:
        calls = []

        def __call__(self, lr, training=False, resolve=False):
            Stub.calls.append(lr.shape[0])
            s = np.asarray(lr, np.float32).sum(axis=3)[:, :16, :16, :]                     # permutation-invariant over T
            return np.round(np.clip(np.repeat(np.repeat(s, 3, 1), 3, 2), 0, 65536))

    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 900, size=(37, 22, 22, 9, 1)).astype(np.float32)
    whole = Stub()(lr)
    Stub.calls.clear()
    by = pb.resolveByBatch(Stub(), lr, batch_size=16)
    assert Stub.calls == [16, 16, 5] and np.array_equal(by, whole)                      # test.py:126-134
    np.random.seed(0)
    avg = pb.resolveBySampleAveraging(Stub(), lr[:4], repeats=5)
    assert np.allclose(avg, whole[:4], atol=0.51)      # the float32 sum over T depends on the frame order only in its last bit, before rounding
    imgs = np.arange(64 * 4).reshape(64, 2, 2, 1).astype(np.float64)
    rec = pb.reconstruct_from_patches(imgs)
    assert rec.shape == (16, 16, 1) and np.array_equal(rec[2:4, 4:6], imgs[1 * 8 + 2])   # row-major: block (i=1, j=2)


def test_cli_auto_precision_falls_back_to_the_dense_engine(monkeypatch):
    """--precision auto (cli._build_model): the preferred tensor-core engine when model creation accepts the graph, the dense fp32 engine
    when it rejects it; an explicit precision is passed through untouched (and its error is NOT swallowed).  No GPU: build_from_config is
    replaced by a stand-in that rejects 64-filter graphs on the row engines like pv_model_create does."""
    import probav_b200 as pb
    from probav_b200 import cli
    from probav_b200._lib import PvError
    calls = []

    def fake_build(config, band="NIR", precision="fp32", **kw):
        calls.append(precision)
        if precision != "fp32" and config["num_filters"] != 32:
            raise PvError("[pv_status -2] the row engine runs 32-filter graphs")
        return ("model", precision)

    monkeypatch.setattr(pb, "build_from_config", fake_build)
    assert cli._build_model({"num_filters": 32}, "NIR", "auto", "tf32x3") == ("model", "tf32x3")
    assert cli._build_model({"num_filters": 64}, "NIR", "auto", "tf32x3") == ("model", "fp32")
    assert calls == ["tf32x3", "tf32x3", "fp32"]
    assert cli._build_model({"num_filters": 32}, "RED", "tf32", "tf32x3") == ("model", "tf32")
    with pytest.raises(PvError):
        cli._build_model({"num_filters": 64}, "NIR", "tf32x3", "tf32x3")
    a, b = cli.train_parser().parse_args([]), cli.test_parser().parse_args([])
    assert a.precision == "auto" and b.precision == "auto"
