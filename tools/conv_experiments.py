"""Which resource bounds a stage of the tensor-core convolution?  Needs a library built with -DGPFQ_CONV_EXPERIMENT
(python tools/conv_experiments.py --build does that; rebuild normally afterwards).  Every experiment removes part of a
k-block's work (results are wrong by construction) and the kernel is timed on ResNet-50 shapes at bs=256:

    1  no lo-plane store by the split warps      (-16 KB shared-memory writes per k-block)
    2  hi*hi products only                       (4 instead of 12 MMAs: a third of the tensor work and operand reads)
    4  drain warps read half their columns       (half the TMEM reads and fp32 adds)
    8  split warps neither load nor store
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if "--build" in sys.argv:
    from quantized_neural_nets_b200 import build
    build.NVCC_FLAGS.append("-DGPFQ_CONV_EXPERIMENT")
    build.build(force=True)
    sys.exit(0)

import torch
from quantized_neural_nets_b200._lib import lib, launch

dev = torch.device("cuda:0")
B = 256
g = torch.Generator(device=dev).manual_seed(0)


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


masks = (0, 1, 2, 4, 8, 3, 6, 12, 15)
print("shape".ljust(22) + "".join(f"{m:>9d}" for m in masks) + "    k-blocks/SM   cycles/k-block (mask 0, 1.9 GHz)")
for (cin, cout, hw) in ((1024, 512, 14), (1024, 256, 14), (512, 256, 28), (512, 128, 28), (256, 64, 56), (256, 128, 56),
                        (64, 256, 56), (64, 64, 56)):
    x = torch.relu(torch.randn(B, cin, hw, hw, device=dev, generator=g))
    w = torch.randn(cout, cin, device=dev, generator=g) * 0.05
    alpha = torch.rand(cout, device=dev, generator=g) + 0.5
    beta = torch.randn(cout, device=dev, generator=g) * 0.1
    out = torch.empty(B, cout, hw, hw, device=dev)
    ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(cout, cin), dtype=torch.uint8, device=dev)

    def fused():
        launch(lib.gpfq_conv1x1_bn_act_f32, x, hw * hw, w, None, alpha, beta, out, B, cin, cout, hw * hw, 0.0,
               float("inf"), ws, ws.numel())

    row = []
    for m in masks:
        os.environ["GPFQ_CONV_EXPERIMENT"] = str(m)
        row.append(t(fused))
    os.environ["GPFQ_CONV_EXPERIMENT"] = "0"
    tiles = B * -(-hw * hw // 128) * -(-cout // 128)
    kb = tiles * -(-cin // 32) / 148
    print(f"{cin:5d}->{cout:5d} @ {hw:3d}".ljust(22) + "".join(f"{v:9.3f}" for v in row) +
          f"    {kb:9.1f}     {row[0] * 1e-3 * 1.9e9 / kb:8.0f}")
