"""Times gpfq_conv1x1_bn_act_f32 on ResNet-50 shapes (bs = 256 by default) for several activation L2-prefetch distances
(GPFQ_CONV_PREFETCH = tiles ahead; 0 = off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_neural_nets_b200._lib import lib, launch

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device=dev).manual_seed(0)
HBM = 6545.3


def t(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


dist = (0, 1, 2, 3, 4, 6)
print(f"B = {B}".ljust(28) + "".join(f"{m:>9d}" for m in dist) + "   best: of HBM")
for (cin, cout, hw, with_res) in ((64, 64, 56, False), (64, 256, 56, False), (64, 256, 56, True), (256, 64, 56, False),
                                  (256, 128, 56, False), (128, 512, 28, True), (512, 128, 28, False), (512, 256, 28, False),
                                  (256, 1024, 14, True), (1024, 256, 14, False), (1024, 512, 14, False)):
    x = torch.relu(torch.randn(B, cin, hw, hw, device=dev, generator=g))
    w = torch.randn(cout, cin, device=dev, generator=g) * 0.05
    alpha = torch.rand(cout, device=dev, generator=g) + 0.5
    beta = torch.randn(cout, device=dev, generator=g) * 0.1
    res = torch.randn(B, cout, hw, hw, device=dev, generator=g) if with_res else None
    out = torch.empty(B, cout, hw, hw, device=dev)
    ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(cout, cin), dtype=torch.uint8, device=dev)

    def fused():
        launch(lib.gpfq_conv1x1_bn_act_f32, x, hw * hw, w, res, alpha, beta, out, B, cin, cout, hw * hw, 0.0,
               float("inf"), ws, ws.numel())

    row = []
    ref = None
    for m in dist:
        os.environ["GPFQ_CONV_PREFETCH"] = str(m)
        row.append(t(fused))
        if ref is None:
            ref = out.clone()
        else:
            assert torch.equal(ref, out), "prefetch distance changed the result"
    bytes_ = 4.0 * B * hw * hw * (cin + cout * (2 if with_res else 1))
    print(f"{cin:5d}->{cout:5d} @ {hw:3d}{' +res' if with_res else ''}".ljust(28) + "".join(f"{v:9.3f}" for v in row) +
          f"   {bytes_ / min(row) / 1e6 / HBM:.2f}", flush=True)
