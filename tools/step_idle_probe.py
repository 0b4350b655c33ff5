"""Where does the GPU wait inside one ResNet-50 quantize_network() step?  Profiles ONE steady-state step with
torch.profiler (CUDA activities), prints wall time vs the union of kernel / memcpy intervals on the device, and the
largest idle gaps with the kernels on either side.  (Profiler overhead inflates the wall time: read the SHARES.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torchvision
from torch.profiler import ProfilerActivity, profile

import quantized_neural_nets_b200 as qb

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
model = torchvision.models.resnet50(weights=None).eval().to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
pool = [torch.randn(B, 3, 224, 224, device=dev, generator=gen) for _ in range(8)]


class Pool:
    def __iter__(self):
        i = 0
        while True:
            yield pool[i % len(pool)], None
            i += 1


def step():
    np.random.seed(0)
    q = qb.QuantizeNeuralNet(model, "resnet50", B, Pool(), 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25, False, dev,
                             solver="auto", fuse_forward=True, pointwise_gemm=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    q.quantize_network()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


for _ in range(3):
    ms = step()
print(f"plain step: {ms:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ms_p = step()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy, cur_end, gaps = 0.0, t0, []
prev = None
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, prev.name if prev else "", e.name))
        busy += en - s
        cur_end = en
        prev = e
    elif en > cur_end:
        busy += en - cur_end
        cur_end = en
        prev = e
span = t1 - t0
print(f"profiled step: {ms_p:.1f} ms; device span {span / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms, idle {100 * (1 - busy / span):.1f} % "
      f"in {len(gaps)} gaps")
agg = {}
for g, a, b in gaps:
    k = (a[:48], b[:48])
    v = agg.setdefault(k, [0, 0.0])
    v[0] += 1
    v[1] += g
print("largest idle totals by (kernel before -> kernel after):")
for (a, b), (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {tot / 1e3:8.2f} ms in {n:5d} gaps   {a} -> {b}")
big = sorted(gaps, reverse=True)[:12]
print("largest single gaps:")
for g, a, b in big:
    print(f"  {g / 1e3:8.3f} ms   {a[:60]} -> {b[:60]}")
