"""Times the batched grouped solver on depthwise shapes (MobileNetV2 at bs=256) against the HBM bound 8*d*m bytes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_neural_nets_b200 import step_algorithm as sa

dev = torch.device("cuda:0")
for groups, dg, m in ((32, 9, 92672), (96, 9, 92672), (144, 9, 23296), (192, 9, 6656), (384, 9, 1792), (960, 9, 768),
                      (96, 25, 23296), (64, 32, 23296)):
    g = torch.Generator(device=dev).manual_seed(0)
    W = torch.randn(groups, dg, device=dev, generator=g) * 0.1
    ld = (m + 3) // 4 * 4
    Xfm = torch.relu(torch.randn(groups * dg, ld, device=dev, generator=g))
    Xqfm = torch.relu(Xfm + 0.02 * torch.randn(groups * dg, ld, device=dev, generator=g))
    X, Xq = Xfm[:, :m].t(), Xqfm[:, :m].t()
    delta = torch.tensor(0.02, device=dev)
    best = 1e9
    for r in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sa.quantize_layer_impl(W, X, Xq, m, 1.16 / 8, 8, 1, None, 0.0, groups, False, dev, solver=sa.GROUPED,
                               return_partials=True, delta=delta)
        b.record()
        torch.cuda.synchronize()
        if r:
            best = min(best, a.elapsed_time(b))
    gb = 8.0 * groups * dg * m / 1e9
    print(f"{groups} groups x {dg} features x {m}: {best:.3f} ms, {gb / best * 1e3:.0f} GB/s of layer input", flush=True)
