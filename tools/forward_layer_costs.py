"""Where the calibration forward's time goes, weighted by the reference's schedule.

The reference re-runs both networks from the image up to layer i for every layer i (quantize_neural_net.py:256-269),
so a module that precedes k quantizable layers is executed 2*k times per quantize_network().  This tool times every
leaf module of the (fused) ResNet-50 forward at bs=256 with CUDA events and prints cost x multiplicity, sorted.

    python tools/forward_layer_costs.py [--no-fuse] [--no-pointwise]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision

import quantized_neural_nets_b200 as qb
from quantized_neural_nets_b200.forward_fusion import fuse_inference_forward, pointwise_convs_as_gemm

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
name = "resnet50"
model = getattr(torchvision.models, name)(weights=None).eval().to(dev)
fuse = "--no-fuse" not in sys.argv
net = fuse_inference_forward(model)[0] if fuse else model
x = torch.randn(256, 3, 224, 224, device=dev)

qlayers = []
qb.extract_layers(model, qlayers)
qset = {id(l): i for i, l in enumerate(qlayers)}

records = {}     # module name -> [events]
order = []


fused_types = ("FusedConvBNAct", "FusedBNAct", "FastMaxPool")
owned = set()       # modules that live inside a fused site: a hook on them would make the site fall back
for n, c in net.named_modules():
    if type(c).__name__ == "FusedConvBNAct":
        owned |= {id(c.conv), id(c.tail), id(c.tail.bn)}
    elif type(c).__name__ == "FusedBNAct":
        owned.add(id(c.bn))
    elif type(c).__name__ == "FastMaxPool":
        owned.add(id(c.pool))
timed = [(n, c) for n, c in net.named_modules()
         if id(c) not in owned and (type(c).__name__ in fused_types or not list(c.children()))]
# a FusedBNAct that is the tail of a FusedConvBNAct is in `owned`; stand-alone ones are timed
handles = []
for n, c in timed:

    def pre(mod, args, n=n):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        records.setdefault(n, []).append([e, None])
        if n not in order:
            order.append(n)

    def post(mod, args, out, n=n):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        records[n][-1][1] = e

    handles.append(c.register_forward_pre_hook(pre))
    handles.append(c.register_forward_hook(post))


def run():
    with torch.no_grad():
        net(x)


import contextlib
ctx = pointwise_convs_as_gemm(model) if "--no-pointwise" not in sys.argv else contextlib.nullcontext()
with ctx:
    for _ in range(3):
        run()
    records.clear()
    reps = 5
    for _ in range(reps):
        run()
    torch.cuda.synchronize()

mods = dict(net.named_modules())
rows = []
seen_q = 0
total_once = 0.0
for n in order:
    ms = sum(a.elapsed_time(b) for a, b in records[n]) / reps
    mod = mods[n]
    # number of quantizable layers that come strictly AFTER this module's position = how many prefix passes run it
    inner = getattr(mod, "conv", None)
    if id(mod) in qset or id(inner) in qset:
        seen_q = qset[id(mod) if id(mod) in qset else id(inner)] + 1
        mult = len(qlayers) - seen_q          # the layer's own forward is interrupted by the hook
    else:
        mult = len(qlayers) - seen_q
    rows.append((n, type(mod).__name__, ms, mult, 2 * ms * mult))
    total_once += ms
step = sum(r[4] for r in rows)
print(f"{name} fused={fuse}: one full forward {total_once:.2f} ms (sum of module times); "
      f"all prefix passes of one quantize_network(): {step:.0f} ms")
print(f"{'module':46s} {'type':14s} {'ms':>8s} {'x passes':>8s} {'ms/step':>9s} {'share':>6s}")
for r in sorted(rows, key=lambda r: -r[4])[:45]:
    print(f"{r[0]:46s} {r[1]:14s} {r[2]:8.3f} {2 * r[3]:8d} {r[4]:9.1f} {r[4] / step:6.3f}")
by_type = {}
for r in rows:
    key = r[1]
    m = mods[r[0]]
    m = getattr(m, "conv", m)
    if isinstance(m, torch.nn.Conv2d):
        key = f"{r[1]} {m.kernel_size[0]}x{m.kernel_size[1]} s{m.stride[0]}"
    by_type[key] = by_type.get(key, 0.0) + r[4]
print("\nby type:")
for k, v in sorted(by_type.items(), key=lambda kv: -kv[1]):
    print(f"  {k:24s} {v:9.1f} ms/step  {v / step:6.3f}")
