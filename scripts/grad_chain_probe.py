"""Where does the data-gradient chain of a tensor-core engine drift from the exact fp32 engine on the same layouts?
python scripts/grad_chain_probe.py [precision] : one forward/backward of the golden B=128 batch on `precision` and on fp32_rows,
then per internal gradient buffer the SYSTEMATIC part of the error (projection coefficient <x, ref> / <ref, ref> - 1) and the
residual rms relative to the rms of the reference."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import probav_b200 as pb
from probav_b200 import _lib, synth
from oracle.wdsr import OracleWDSR, init_params

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
FULL = dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=12, expRate=8, decayRate=0.8, numImgLR=9, patchSizeLR=16, isGrayScale=True)
om = OracleWDSR(8075.2045, 3160.7272, 6, **FULL)
p = init_params(om.specs, seed=100, dtype=torch.float64)
lr, hr, mask = synth.make_batch(128, seed=101, hr_zero_under_mask=False)
lr, hr, mask = lr[:B], hr[:B], mask[:B]
bufs = {}
import tempfile
for pr in (prec, "fp32_rows"):
    m = pb.WDSRConv3D("superResolutionNet", "NIR", 8075.2045, 3160.7272, 6).build(**FULL, precision=pr)
    m.set_weights({k: v.numpy().astype(np.float32) for k, v in p.items()})
    L = pb.Losses((48, 48, 1))
    d = tempfile.mkdtemp()
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/c", d + "/l")
    t.forward_backward(lr, hr, mask)
    # only the LAST values of the ping-pong gradient buffers survive; the stored activations all do
    names = ["g_U", "g_Go3", "g_Go2", "g_Go1", "g_Gi1", "g_D", "g_a0", "g_a1", "a0", "a6", "a12", "D0", "D11", "Go3", "U"]
    out = {}
    for n in names:
        ln = _lib.lib().pv_debug_read_buffer(m._h, n.encode(), 1, None, 0)
        if ln <= 0:
            continue
        a = np.empty(ln, np.float32)
        _lib.lib().pv_debug_read_buffer(m._h, n.encode(), 1, a.ctypes.data_as(C.c_void_p), ln)
        out[n] = a
    bufs[pr] = out
    m.close()
for n, ref in bufs["fp32_rows"].items():
    x = bufs[prec].get(n)
    if x is None or x.shape != ref.shape:
        continue
    r64, x64 = ref.astype(np.float64), x.astype(np.float64)
    den = float((r64 * r64).sum())
    if den == 0:
        continue
    coef = float((x64 * r64).sum()) / den
    resid = x64 - coef * r64
    print(f"{n:8s} scale error {coef - 1:+.3e}   residual rms / ref rms {np.sqrt((resid ** 2).mean() / (r64 ** 2).mean()):.3e}   max|err|/max|ref| {np.abs(x64 - r64).max() / np.abs(r64).max():.3e}")
