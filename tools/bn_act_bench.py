"""Times gpfq_bn_act_f32 on ResNet-50 activation shapes at bs=256 against HBM bytes (read x (+ residual), write out)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_neural_nets_b200.forward_fusion import FusedBNAct

dev = torch.device("cuda:0")
for (B, C, H, res) in ((256, 64, 112, False), (256, 64, 56, False), (256, 256, 56, True), (256, 128, 28, False),
                       (256, 512, 28, True), (256, 256, 14, False), (256, 1024, 14, True), (256, 512, 7, False),
                       (256, 2048, 7, True)):
    bn = torch.nn.BatchNorm2d(C).eval().to(dev)
    f = FusedBNAct(bn, 0.0, float("inf"))
    x = torch.randn(B, C, H, H, device=dev)
    r = torch.randn(B, C, H, H, device=dev) if res else None
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    best = 1e9
    for it in range(5):
        flush.zero_()                                   # 256 MB write: evicts x / r from the 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        y = f(x, r)
        b.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, a.elapsed_time(b))
    gb = x.numel() * 4 * (3 if res else 2) / 1e9
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    flush.zero_(); t0.record()
    z = torch.relu(bn(x) + r) if res else torch.relu(bn(x))
    t1.record(); torch.cuda.synchronize()
    print(f"({B},{C},{H},{H}) res={res}: fused {best*1e3:.0f} us = {gb/best*1e3:.0f} GB/s; torch bn(+add)+relu {t0.elapsed_time(t1)*1e3:.0f} us",
          flush=True)
