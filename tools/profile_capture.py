"""One conv-input capture (unfold with stride = kernel + Bernoulli-subsampled row gather + transpose: im2col_gather_kernel),
repeated -- the command line for `ncu -k regex:im2col_gather_kernel`:
    python tools/profile_capture.py C H K [B]      (default: 64 56 1 256 = ResNet-50 layer1 1x1 layers, m = 200 960)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quantized_neural_nets_b200 import SaveInputConv2d

C, H, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 56, 1)))
B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
dev = torch.device("cuda:0")
x = torch.randn(B, C, H, H, device=dev)
np.random.seed(0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    s = SaveInputConv2d(kernel_size=K, dilation=1, padding=K // 2, stride=1, groups=1, retain_rate=0.25)
    if it == 3:
        a.record()
    X = s.capture(x)
b.record()
torch.cuda.synchronize()
m, d = X.shape
print(f"capture of a ({B}, {C}, {H}, {H}) input, {K}x{K} patches: X is {m} x {d}; {a.elapsed_time(b) / 3:.3f} ms per capture "
      f"(host index draw included); algorithmic bytes read+written {8e-6 * m * d:.1f} MB")
