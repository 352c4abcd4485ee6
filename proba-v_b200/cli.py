"""Command-line entry points with the reference's flags: `train.py --cfg --band --modelType` (reference train.py:26-32,35-110)
and `test.py --cfg --band --totest` (reference test.py:25-31,34-100).  Same cfg file, same input files
(`<preprocessing_out>/augmentedPatchesDir/TRAIN[VAL]patches{LR,HR}_<band>.npy` masked arrays for training,
`<preprocessing_out>/resolverDir/<totest>patchesLR_<band>.npy` for prediction), same outputs
(`<model_out>/ckpt_<cfg>/<band>/ckpt-N.*` TensorFlow checkpoints, `<model_out>/logs_<cfg>/<band>/events.out.tfevents.*`,
`<test_out>_<cfg>/imgsetNNNN.png` 16-bit PNGs).  `--synthetic N` substitutes seeded synthetic patches / scenes for the
.npy files (there is no dataset in this environment).  Only the patchNet model type is built; the reference's fusionNet
branch (train.py:113-187) reads hard-coded paths of its author's machine and is out of scope."""
from __future__ import annotations

import argparse
import logging
import os
import struct
import zlib

import numpy as np

logger = logging.getLogger("probav_b200")


def write_png16(path: str, img: np.ndarray):
    """skimage.io.imsave(path, img.astype(np.uint16)) for a 2-D image (test.py:99): 16-bit grayscale PNG, no extra dependency."""
    a = np.ascontiguousarray(np.asarray(img).astype(np.uint16))
    if a.ndim != 2:
        raise ValueError("write_png16 expects a 2-D image")
    h, w = a.shape
    raw = b"".join(b"\x00" + a[y].astype(">u2").tobytes() for y in range(h))        # filter type 0 per scan line

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 0, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def read_png16(path: str) -> np.ndarray:
    """Inverse of write_png16 (filter type 0 only) -- used by the tests."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", 0
    h = 0
    while pos < len(b):
        (n,) = struct.unpack(">I", b[pos:pos + 4])
        tag, data = b[pos + 4:pos + 8], b[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", data[:10])
            assert (depth, ctype) == (16, 0)
        elif tag == b"IDAT":
            idat += data
        pos += 12 + n
    raw = zlib.decompress(idat)
    rows = np.frombuffer(raw, np.uint8).reshape(h, 1 + 2 * w)
    assert not rows[:, 0].any()
    return rows[:, 1:].copy().view(">u2").astype(np.uint16).reshape(h, w)


def _band_stats(band: str):
    from .synth import BAND_STATS
    return BAND_STATS["NIR" if band == "NIR" else "RED"]          # train.py:47-52 / test.py:40-45


def _dirs(config: dict, cfg_path: str, band: str):
    basename = os.path.basename(cfg_path).split(".")[0]
    return (basename, os.path.join(config["model_out"], f"ckpt_{basename}", band),
            os.path.join(config["model_out"], f"logs_{basename}", band))


def _build_model(config, band, precision, preferred):
    """--precision auto: the tensor-core engine `preferred` when it runs this graph (32 filters, 7 / 9 / 13 LR frames), else the dense
    fp32 CUDA-core engine, which builds every graph the reference can (both are GPU engines: there is no CPU path)."""
    from . import build_from_config
    from ._lib import PvError
    if precision != "auto":
        return build_from_config(config, band=band, precision=precision)
    try:
        return build_from_config(config, band=band, precision=preferred)
    except (PvError, ValueError) as e:
        logger.info(f"[ INFO ] {preferred} engine does not run this graph ({e}); using the dense fp32 engine")
        return build_from_config(config, band=band, precision="fp32")


def train_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg/yourcfg.cfg", type=str)
    ap.add_argument("--band", type=str, default="NIR")
    ap.add_argument("--modelType", type=str, default="patchNet")
    ap.add_argument("--synthetic", type=int, default=0, help="train on N seeded synthetic patches instead of the .npy files")
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--eval-step", type=int, default=1000)
    ap.add_argument("--precision", default="auto", choices=["auto", "tf32", "tf32x3", "fp32", "fp32_rows"],
                    help="auto = tf32x3 (error-compensated tensor cores: gradients inside the 1e-3 bar) where the row engine runs the graph, else fp32")
    return ap


def train_main(argv=None):
    from . import Losses, ModelTrainer, build_from_config, loss_from_config, optimizers, parallel, parseConfig, synth
    opt = train_parser().parse_args(argv)
    if opt.modelType != "patchNet":
        raise SystemExit("only --modelType patchNet is built (the reference's fusionNet branch is out of scope)")
    config = parseConfig(opt.cfg)
    parallel.init_from_env()
    logger.info("[ INFO ] Loading data...")
    if opt.synthetic:
        nval = max(config["batch_size"], opt.synthetic // 5)
        X_train, y_train, y_train_mask = synth.make_batch(opt.synthetic, T=config["num_low_res_imgs"], patch=config["patch_size"],
                                                          scale=config["scale"], max_shift=config["max_shift"], band=opt.band, seed=1)
        X_val, y_val, y_val_mask = synth.make_batch(nval, T=config["num_low_res_imgs"], patch=config["patch_size"],
                                                    scale=config["scale"], max_shift=config["max_shift"], band=opt.band, seed=2)
    else:
        dataDir = os.path.join(config["preprocessing_out"], "augmentedPatchesDir")
        ld = lambda n: np.load(os.path.join(dataDir, f"{n}_{opt.band}.npy"), allow_pickle=True)      # noqa: E731
        X_train, X_val, y_train, y_val = ld("TRAINpatchesLR"), ld("TRAINVALpatchesLR"), ld("TRAINpatchesHR"), ld("TRAINVALpatchesHR")
        y_train_mask, y_val_mask = ~np.ma.getmaskarray(y_train), ~np.ma.getmaskarray(y_val)          # True = clear (train.py:43-44)
        X_train, X_val, y_train, y_val = (np.asarray(np.ma.getdata(a), np.float32) for a in (X_train, X_val, y_train, y_val))
    logger.info("[ INFO ] Building model...")
    model = _build_model(config, opt.band, opt.precision, "tf32x3")
    target = config["scale"] * config["patch_size"]
    loss = Losses(targetShape=(target, target, 1))
    basename, ckptDir, logDir = _dirs(config, opt.cfg, opt.band)
    trainer = ModelTrainer(model=model, loss=loss_from_config(loss, config["loss"]), metric=loss.shiftCompensatedcPSNR,
                           optimizer=optimizers.from_config(config["optimizer"], config["learning_rate"]),
                           ckptDir=ckptDir, logDir=logDir, evalStep=opt.eval_step)
    if parallel.world()[1] > 1:
        parallel.broadcast_(model.param_arena(), 0)
    trainer.fitTrainData(X_train, [y_train, y_train_mask], config["batch_size"] * parallel.world()[1], config["epochs"],
                         [X_val, y_val, y_val_mask], saveBestOnly=False, initEpoch=0, maxSteps=opt.max_steps)
    if opt.max_steps:                     # a bounded run still leaves a checkpoint behind, like an evaluation step would
        trainer.save()
    trainer.close()
    logger.info(f"[ SUCCESS ] Model checkpoint can be found in {ckptDir}.")
    logger.info(f"[ SUCCESS ] Model logs can be found in {logDir}.")
    return ckptDir, logDir


def test_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg/FINAL.cfg", type=str)
    ap.add_argument("--band", type=str, default="RED")
    ap.add_argument("--totest", type=str, default="TEST")
    ap.add_argument("--synthetic", type=int, default=0, help="predict N seeded synthetic 128x128 scenes instead of the .npy file")
    ap.add_argument("--precision", default="auto", choices=["auto", "tf32", "tf32x3", "fp32", "fp32_rows"],
                    help="auto = tf32 (single-pass tensor cores: SR inside the 1e-3 bar) where the row engine runs the graph, else fp32")
    return ap


def test_main(argv=None):
    from . import build_from_config, evaluate, parseConfig, synth
    opt = test_parser().parse_args(argv)
    config = parseConfig(opt.cfg)
    model = _build_model(config, opt.band, opt.precision, "tf32")
    basename, ckptDir, _ = _dirs(config, opt.cfg, opt.band)
    try:
        info = model.restore_checkpoint(ckptDir)          # ckpt.restore(ckptMngr.latest_checkpoint), test.py:58-67
        logger.info(f"[ INFO ] Restored {ckptDir} (step {info['step']}).")
    except FileNotFoundError:
        logger.info(f"[ INFO ] No checkpoint in {ckptDir}: predicting with the initial weights (what TF's restore(None) does).")
    logger.info("[ INFO ] Generating predictions...")
    if opt.synthetic:
        scenes, _, _ = synth.make_scene(opt.synthetic, T=config["num_low_res_imgs"], band=opt.band, seed=3)
        y_preds = list(model.predict_from_scenes(scenes).astype(np.float64))      # patching + stitching on the device
    else:
        dataDir = os.path.join(config["preprocessing_out"], "resolverDir")
        patchLR = np.load(os.path.join(dataDir, f"{opt.totest}patchesLR_{opt.band}.npy"), allow_pickle=True)
        patchLR = np.asarray(np.ma.getdata(patchLR), np.float32).transpose((0, 1, 4, 5, 2, 3))         # test.py:38
        y_preds = evaluate(model, patchLR)
    band = opt.band.upper()
    toOmit = []
    if os.path.exists(f"removedTrainSets{band}.txt"):
        toOmit = [int(float(x.strip())) for x in open(f"removedTrainSets{band}.txt") if x.strip()]
    if opt.totest == "TEST":
        outDir, i = config["test_out"] + f"_{basename}", (1306 if band == "NIR" else 1160)          # test.py:80-85
    else:
        outDir, i = config["train_out"] + f"_{basename}", (594 if band == "NIR" else 0)
    os.makedirs(outDir, exist_ok=True)
    logger.info(f"[ SAVE ] Saving predicted images to {outDir}...")
    for img in y_preds:
        while i in toOmit:
            i += 1
        write_png16(os.path.join(outDir, f"imgset{'%04d' % i}.png"), img[:, :, 0])
        i += 1
    return outDir
