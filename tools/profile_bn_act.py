"""One gpfq_bn_act_f32 launch on a (256, 256, 56, 56) activation with a residual (for ncu --set full)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_neural_nets_b200.forward_fusion import FusedBNAct

dev = torch.device("cuda:0")
bn = torch.nn.BatchNorm2d(256).eval().to(dev)
f = FusedBNAct(bn, 0.0, float("inf"))
x = torch.randn(256, 256, 56, 56, device=dev)
r = torch.randn(256, 256, 56, 56, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for _ in range(3):
    flush.zero_()
    y = f(x, r)
torch.cuda.synchronize()
print("ok", float(y.sum()))
