"""Optimizer selectors with the Keras constructor surface the reference uses (train.py:76-83):
Adam / Nadam / SGD(learning_rate=...).  They only carry hyper-parameters; the update itself is the fused
arena kernel behind pv_apply_gradients (csrc/glue.cu, csrc/engine.cu trainer_apply)."""


class _Opt:
    kind = "sgd"

    def __init__(self, learning_rate=0.001):
        self.learning_rate = float(learning_rate)


class Nadam(_Opt):
    kind = "nadam"


class Adam(_Opt):
    kind = "adam"


class SGD(_Opt):
    kind = "sgd"

    def __init__(self, learning_rate=0.01):
        super().__init__(learning_rate)


def from_config(name: str, learning_rate: float) -> _Opt:
    """train.py:77-83: 'adam' | 'nadam' | anything else -> SGD."""
    if name == "adam":
        return Adam(learning_rate)
    if name == "nadam":
        return Nadam(learning_rate)
    return SGD(learning_rate)
