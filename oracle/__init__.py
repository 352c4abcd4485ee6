"""CPU oracle for the PROBA-V 3D-WDSR hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch fp64/fp32 + NumPy + one plain-C file)
of the reference algorithm for the train/infer hot path of mmbajo/PROBA-V:

    models/modelsTF.py:7-203     -> oracle/wdsr.py
    models/loss.py:8-97,126-187,219-238 + utils/utils.py:42-44 -> oracle/losses.py, oracle/shift_loss.c
    models/trainClass.py:124-143 -> oracle/step.py
    train.py:76-83 (Keras Nadam/Adam/SGD) -> oracle/optim.py
    test.py:103-160, models/testClass.py:24-39, utils/dataGenerator.py:108-121 -> oracle/step.py

PARITY UNPINNED: the reference is TensorFlow 2.x + tensorflow-addons, neither of
which is installed in this image (no network), and the reference ships no tests,
golden vectors or trained weights.  The oracle therefore restates the published
TF/TFA semantics and is pinned only by hand-derived
known-answer cases and cross-checks against torch equivalents (tests/test_oracle.py).
(The on-disk formats are a different matter: the checkpoint and event-file code of proba-v_b200/
is pinned byte for byte by the reference's own files, tests/test_tfckpt.py and tests/test_tbevents.py.)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (proba-v_b200/) never does.
"""
