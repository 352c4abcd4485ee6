"""CPU: the C-ABI library loads and exports every symbol include/probav_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "probav_b200.h")
LIB = os.path.join(ROOT, "proba-v_b200", "libprobav_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("pv_model_create", "pv_forward", "pv_resolve", "pv_shift_loss", "pv_train_step", "pv_eval_step",
                 "pv_train_forward_backward", "pv_apply_gradients", "pv_predict_scenes_host", "pv_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"libprobav_b200.so does not export {missing}"
    lib.pv_abi_version.restype = ctypes.c_int
    assert lib.pv_abi_version() == 1


def test_python_binding_covers_every_declared_symbol():
    import probav_b200  # noqa: F401
    from probav_b200._lib import _SIGS
    assert sorted(_SIGS) == declared_symbols()


def test_no_cpu_fallback_without_device():
    """On a box without a B200 the product path must fail loudly, never fall back to the oracle/CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import probav_b200 as pb
    from probav_b200._lib import PvError
    with pytest.raises(PvError):
        pb.WDSRConv3D("n", "NIR", 8075.2, 3160.7, 6).build(3, 32, (3, 3, 3), 2, 8, 0.8, 9, 16, True)
    import numpy as np
    z = np.zeros((1, 48, 48, 1), np.float32)
    with pytest.raises(PvError):
        pb.Losses((48, 48, 1)).shiftCompensatedL1Loss(z, np.ones_like(z, bool), z)


def test_bad_cfg_is_rejected_like_the_reference_graph_would():
    # num_low_res_imgs=12 has no reducer branch (modelsTF.py:62-69); max_shift != 6 breaks the Reshape (:71)
    import probav_b200 as pb
    with pytest.raises(ValueError):
        pb.WDSRConv3D("n", "NIR", 8075.2, 3160.7, 6).build(3, 32, (3, 3, 3), 2, 8, 0.8, 12, 16, True)
    with pytest.raises(ValueError):
        pb.WDSRConv3D("n", "NIR", 8075.2, 3160.7, 4).build(3, 32, (3, 3, 3), 2, 8, 0.8, 9, 16, True)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch types) and a C program must link against
    the library and call into it (pv_abi_version / pv_device_count / pv_last_error need no GPU)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    subprocess.check_call([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", "-pedantic", HEADER])
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "probav_b200.h"\n'
                   'int main(void) { pv_model* m = 0; pv_cfg c = {0}; int rc = pv_model_create(&c, 0, &m);\n'
                   '  printf("%d %d %d\\n", pv_abi_version(), rc != 0, pv_last_error()[0] != 0); return 0; }\n')
    exe = tmp_path / "t"
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe),
                           "-L", os.path.dirname(LIB), "-lprobav_b200", "-Wl,-rpath," + os.path.dirname(LIB)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout.split()
    assert out == ["1", "1", "1"]          # ABI version 1; an all-zero cfg is rejected with a message, no crash
