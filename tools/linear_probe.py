#!/usr/bin/env python
"""Developer probe: time of the head linear (b2me_linear_small, 1024 -> 3 with arg-max) on a 32-frame-sized input."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "markerless-robot-camera-calibration_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from MinkowskiEngine._lib import lib, ptr, stream, check, BF16  # noqa: E402

V, Cin, Cout = 8918163, 1024, 3
x = torch.randn(V, Cin, device="cuda", dtype=torch.bfloat16)
Wt = torch.randn(Cout, Cin, device="cuda") / 32
bias = torch.zeros(Cout, device="cuda")
logits = torch.empty(V, Cout, device="cuda")
am = torch.empty(V, dtype=torch.uint8, device="cuda")


def run():
    check(lib.b2me_linear_small(ptr(x), BF16, V, Cin, ptr(Wt), ptr(bias), Cout, ptr(logits), ptr(am), stream()))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{os.environ.get('B2ME_LIB_PATH', 'product lib')}: {ms:.3f} ms, {V * Cin * 2 / ms / 1e9:.2f} TB/s")
