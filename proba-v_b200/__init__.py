"""probav_b200 -- B200-native (sm_100a) 3D-WDSR train/infer hot path of mmbajo/PROBA-V behind the reference's API.

    from probav_b200 import parseConfig, WDSRConv3D, Losses, ModelTrainer, Enhancer

All compute lives in libprobav_b200.so (include/probav_b200.h); there is no CPU fallback."""
from .parseConfig import parseConfig                      # noqa: F401
from .models import WDSRConv3D, WDSRModel, build_from_config   # noqa: F401
from .loss import Losses, loss_from_config                # noqa: F401
from .trainClass import ModelTrainer                      # noqa: F401
from .testClass import (Enhancer, calcRelativePSNR, evaluate, resolve, resolveByBatch, resolveBySampleAveraging,  # noqa: F401
                        reconstruct_from_patches)
from .optimizers import Adam, Nadam, SGD                  # noqa: F401
from . import optimizers, parallel, synth                 # noqa: F401

__version__ = "0.1.0"
