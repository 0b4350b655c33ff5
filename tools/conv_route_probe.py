"""Stem / stride-2 convolutions of ResNet-50 at bs=256: cuDNN conv2d + bn_act against the patch-matrix route
(gpfq_conv_patches_f32 + gpfq_conv1x1_bn_act_f32), with the two kernels of the route timed separately."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from quantized_neural_nets_b200._lib import lib, launch

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True


def t(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


B = 256
g = torch.Generator(device=dev).manual_seed(0)
B = 256
g = torch.Generator(device=dev).manual_seed(0)
for (cin, cout, h, k, s, p) in ((3, 64, 224, 7, 2, 3), (128, 128, 56, 3, 2, 1), (256, 512, 56, 1, 2, 0), (256, 256, 28, 3, 2, 1),
                                (512, 1024, 28, 1, 2, 0), (512, 512, 14, 3, 2, 1), (1024, 2048, 14, 1, 2, 0),
                                (512, 2048, 7, 1, 1, 0), (2048, 512, 7, 1, 1, 0)):
    x = torch.relu(torch.randn(B, cin, h, h, device=dev, generator=g))
    w = torch.randn(cout, cin, k, k, device=dev, generator=g) * 0.05
    alpha = torch.rand(cout, device=dev, generator=g) + 0.5
    beta = torch.randn(cout, device=dev, generator=g) * 0.1
    ho = (h + 2 * p - (k - 1) - 1) // s + 1
    hw = ho * ho
    ld = (hw + 3) // 4 * 4
    ck = cin * k * k
    out = torch.empty(B, cout, ho, ho, device=dev)
    tmp = torch.empty_like(out)
    patches = torch.empty(B, ck, ld, device=dev)
    ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(cout, ck), dtype=torch.uint8, device=dev)

    def cudnn():
        y = F.conv2d(x, w, stride=s, padding=p)
        launch(lib.gpfq_bn_act_f32, y, None, alpha, beta, tmp, B * cout, cout, hw, 0.0, float("inf"))

    def gather():
        launch(lib.gpfq_conv_patches_f32, x, B, cin, h, h, k, k, s, s, p, p, 1, 1, patches, ld)

    def gemm():
        launch(lib.gpfq_conv1x1_bn_act_f32, patches, ld, w, None, alpha, beta, out, B, ck, cout, hw, 0.0, float("inf"), ws,
               ws.numel())

    with torch.no_grad():
        tc, tg, tm = t(cudnn), t(gather), t(gemm)
        cudnn(); gather(); gemm()
        err = ((out - tmp).norm() / tmp.norm()).item()
    print(f"{cin:5d}->{cout:5d} {k}x{k} s{s} @{h:3d}: cuDNN+bn_act {tc:.3f} ms | patches {tg:.3f} ms ({4e-6 * B * ck * ld / tg:.0f} GB/s "
          f"written) + gemm {tm:.3f} ms = {tg + tm:.3f} ms   rel diff {err:.1e}", flush=True)
