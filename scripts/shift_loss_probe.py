"""Runs the shift-loss kernel alone (for ncu captures and the HBM roofline view): python scripts/shift_loss_probe.py [B] [kind] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import probav_b200 as pb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = sys.argv[2] if len(sys.argv) > 2 else "l1"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(0)
hr = torch.round(torch.rand(B, 48, 48, 1, device="cuda", generator=g) * 4000 + 6000)
sr = hr.roll((1, -2), (1, 2)) + torch.randn(B, 48, 48, 1, device="cuda", generator=g) * 40
mask = torch.rand(B, 48, 48, 1, device="cuda", generator=g) > 0.08
L = pb.Losses((48, 48, 1))
for _ in range(reps):
    L.evaluate(kind, hr, mask, sr, want_grad=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    L.evaluate(kind, hr, mask, sr, want_grad=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"shift loss {kind} + dSR, B={B}: {ms:.4f} ms per call, {B * 2304 * 13 / ms / 1e6:.1f} GB/s algorithmic")
