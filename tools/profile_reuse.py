"""Host-side profile of one calibration='reuse' quantize_network() of ResNet-50 (where does the CPU spend its time
once the forward passes are no longer O(L^2)).  usage: python tools/profile_reuse.py [model] [batch]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torchvision
import quantized_neural_nets_b200 as qb

name = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = getattr(torchvision.models, name)(weights=None).eval().to(dev)
images = torch.randn(batch, 3, 224, 224, device=dev)


def step(profile=False):
    np.random.seed(0)
    qnn = qb.QuantizeNeuralNet(model, name, batch, [(images, None)], 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25, False,
                               dev, calibration="reuse", solver="auto", profile=profile)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    qnn.quantize_network()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    return qnn, t_cpu, time.perf_counter() - t0


for _ in range(3):
    step()
qnn, t_cpu, t_all = step()
print(f"reuse step: host returns after {t_cpu*1e3:.1f} ms, device done after {t_all*1e3:.1f} ms")
qnn, _, _ = step(profile=True)
totals, _ = qnn.phase_times_ms()
print({k: round(v, 1) for k, v in totals.items()})
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
