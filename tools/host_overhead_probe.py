"""Host time per call of the fused forward modules (tiny tensors, so the GPU is never the limit) and of one full
ResNet-50 forward at small batch: is the calibration forward host-bound at B/8 images per GPU?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision

from quantized_neural_nets_b200.forward_fusion import FusedConvBNAct, FusedBNAct, fuse_inference_forward

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
conv = torch.nn.Conv2d(64, 64, 1, bias=False).to(dev)
bn = torch.nn.BatchNorm2d(64).eval().to(dev)
x = torch.randn(1, 64, 8, 8, device=dev)
for name, mod, args in (("FusedConvBNAct", FusedConvBNAct(conv, bn, 0.0, float("inf")), (x,)),
                        ("FusedBNAct", FusedBNAct(bn, 0.0, float("inf")), (x,)), ("nn.Conv2d (cuDNN)", conv, (x,))):
    with torch.no_grad():
        for _ in range(50):
            mod(*args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2000):
            mod(*args)
        dt = time.perf_counter() - t0
        torch.cuda.synchronize()
    print(f"{name:20s} {dt / 2000 * 1e6:7.1f} us of host time per call")
torch.manual_seed(0)
model = torchvision.models.resnet50(weights=None).eval().to(dev)
gm, _ = fuse_inference_forward(model)
for B in (32, 64, 256):
    xb = torch.randn(B, 3, 224, 224, device=dev)
    with torch.no_grad():
        for _ in range(3):
            gm(xb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            gm(xb)
        host = (time.perf_counter() - t0) / 10          # enqueue time (asynchronous)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            gm(xb)
        b.record()
        torch.cuda.synchronize()
    print(f"fused ResNet-50 forward B={B:3d}: host enqueue {host * 1e3:6.2f} ms, device {a.elapsed_time(b) / 10:6.2f} ms per pass")
