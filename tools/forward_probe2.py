"""Full ResNet-50 forward at bs=256: plain / elementwise-fused / fused + 1x1 convolutions as SGEMM; kernel table of the last."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision
from quantized_neural_nets_b200.forward_fusion import fuse_inference_forward, pointwise_convs_as_gemm

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = torchvision.models.resnet50(weights=None).eval().to(dev)
gm, _ = fuse_inference_forward(model)
x = torch.randn(256, 3, 224, 224, device=dev)


def t(fn, n=5):
    with torch.no_grad():
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print(f"plain {t(lambda: model(x)):.1f} ms   fused {t(lambda: gm(x)):.1f} ms", flush=True)
with pointwise_convs_as_gemm(model):
    print(f"fused + pointwise {t(lambda: gm(x)):.1f} ms   plain + pointwise {t(lambda: model(x)):.1f} ms", flush=True)
    from torch.profiler import profile, ProfilerActivity
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
        gm(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
