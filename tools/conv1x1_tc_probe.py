"""Time and check gpfq_conv1x1_bn_act_f32 (tcgen05 split-TF32, fused BN/residual/ReLU) on the ResNet-50 1x1 shapes at
bs=256 against cuDNN conv2d + the separate bn_act pass and against a float64 reference."""
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
HBM = 6545.3


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


B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tot_f = tot_c = 0.0
g = torch.Generator(device=dev).manual_seed(0)
for (cin, cout, hw, count, with_res) in ((64, 64, 56, 1, False), (64, 256, 56, 1, False), (64, 256, 56, 3, True),
                                         (256, 64, 56, 2, False), (256, 128, 56, 1, False), (128, 512, 28, 4, True),
                                         (512, 128, 28, 3, False), (512, 256, 28, 1, False), (256, 1024, 14, 6, True),
                                         (1024, 256, 14, 5, False), (1024, 512, 14, 1, False)):
    x = torch.relu(torch.randn(B, cin, hw, hw, device=dev, generator=g))
    w = torch.randn(cout, cin, device=dev, generator=g) * 0.05
    alpha = torch.rand(cout, device=dev, generator=g) + 0.5
    beta = torch.randn(cout, device=dev, generator=g) * 0.1
    res = torch.randn(B, cout, hw, hw, device=dev, generator=g) if with_res else None
    out = torch.empty(B, cout, hw, hw, device=dev)
    ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(cout, cin), dtype=torch.uint8, device=dev)

    def fused():
        launch(lib.gpfq_conv1x1_bn_act_f32, x, hw * hw, w, res, alpha, beta, out, B, cin, cout, hw * hw, 0.0, float("inf"), ws, ws.numel())

    tmp = torch.empty_like(out)

    def unfused():
        y = F.conv2d(x, w.view(cout, cin, 1, 1))
        launch(lib.gpfq_bn_act_f32, y, res, alpha, beta, tmp, B * cout, cout, hw * hw, 0.0, float("inf"))

    with torch.no_grad():
        tf, tu = t(fused), t(unfused)
        fused()
        nb = min(B, 4)
        ref = torch.einsum("nc,bchw->bnhw", w.double(), x[:nb].double()) * alpha.double()[None, :, None, None] \
            + beta.double()[None, :, None, None]
        if res is not None:
            ref = ref + res[:nb].double()
        ref = ref.clamp(min=0)
        mag = torch.einsum("nc,bchw->bnhw", w.double().abs(), x[:nb].double().abs()) * alpha.double()[None, :, None, None]
        err = ((out[:nb].double() - ref).abs() / (mag + 1e-30)).max().item()
        l2 = ((out[:nb].double() - ref).norm() / ref.norm()).item()
    bytes_ = 4.0 * B * hw * hw * (cin + cout * (2 if with_res else 1))
    fl = 2.0 * B * hw * hw * cin * cout
    tot_f += tf * count
    tot_c += tu * count
    print(f"{cin:5d}->{cout:5d} @{hw:3d}{' +res' if with_res else '     '}: fused {tf:.3f} ms ({bytes_ / tf / 1e6:7.0f} GB/s = "
          f"{bytes_ / tf / 1e6 / HBM:.2f} of HBM, {fl / tf / 1e9:6.1f} TF/s)   cuDNN conv + bn_act {tu:.3f} ms   "
          f"worst |err| / sum|terms| {err:.1e}, rel L2 {l2:.1e}", flush=True)
print(f"per full ResNet-50 forward (stride-1 1x1 layers with HW % 4 == 0): fused {tot_f:.2f} ms, cuDNN + bn_act {tot_c:.2f} ms")
