"""cfg-file reader with the reference's grammar (reference utils/parseConfig.py:5-82).

Sections [Directories] [Train] [Net] [Preprocessing] are flattened into one dict; values are typed by the
section they sit in; any key outside the reference's whitelist raises AssertionError like the reference does
(parseConfig.py:74).  Written from the grammar, not from the reference code.
"""
from __future__ import annotations

import os
from typing import Dict

SUPPORTED_FIELDS = {
    "raw_data", "preprocessing_out", "model_out", "train_out", "test_out",
    "batch_size", "epochs", "learning_rate", "optimizer", "split", "loss",
    "num_res_blocks", "num_low_res_imgs", "num_low_res_imgs_pre", "scale", "num_filters", "kernel_size",
    "exp_rate", "decay_rate", "is_grayscale",
    "max_shift", "patch_size", "patch_stride", "low_res_patch_thresholds", "low_res_threshold",
    "high_res_threshold", "num_low_res_permute", "to_flip", "to_rotate", "ckpt",
}


def _int_list(s):
    return [int(t) for t in s.split(",")]


def _float_list(s):
    return [float(t) for t in s.split(",")]


def _flag(s):
    return bool(int(s.strip()))


def _convert(section: str, key: str, raw: str):
    """Typing rules, first matching substring wins (parseConfig.py:31-59)."""
    if section == "Preprocessing":
        rules = (("ckpt", _int_list), ("low_res_patch_thresholds", _float_list), ("low_res_threshold", float),
                 ("high_res_threshold", float), ("to_flip", _flag), ("to_rotate", _flag))
        default = int
    elif section == "Net":
        rules = (("decay_rate", float), ("is_grayscale", _flag))
        default = int
    elif section == "Train":
        rules = (("learning_rate", float), ("split", float), ("optimizer", str.strip), ("loss", str.strip))
        default = int
    else:
        return raw.strip()
    for needle, fn in rules:
        if needle in key:
            return fn(raw.strip()) if fn in (float, int) else fn(raw)
    return default(raw.strip())


def resolve_path(path: str) -> str:
    if not path.endswith(".cfg"):
        path += ".cfg"
    alt = os.path.join("cfg", path)
    if not os.path.exists(path) and os.path.exists(alt):
        path = alt
    return path


def parseConfig(path: str) -> Dict:
    path = resolve_path(path)
    with open(path, "r") as fh:
        text = fh.read()
    sections = []                       # [(name, {key: value})]
    for line in text.split("\n"):
        if not line or line.startswith("#"):
            continue
        line = line.strip()
        if not line:
            continue
        if line.startswith("["):
            sections.append((line[1:-1].strip(), {}))
            continue
        if not sections:
            raise ValueError(f"{path}: key/value line before any [Section]: {line!r}")
        key, raw = line.split("=")      # exactly one '=' per line, as in the reference
        name, body = sections[-1]
        key = key.strip()
        body[key] = _convert(name, key, raw)
    # whitelist check skips the first section (the reference iterates moduleDefs[1:])
    seen = []
    for _name, body in sections[1:]:
        for k in body:
            if k not in seen:
                seen.append(k)
    unsupported = [k for k in seen if k not in SUPPORTED_FIELDS and k != "type"]
    assert not any(unsupported), "Unsupported fields {} in {}".format(unsupported, path)
    config: Dict = {}
    for _name, body in sections:
        config.update(body)
    return config
