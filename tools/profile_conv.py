"""One fused 1x1 convolution shape, repeated -- the command line for `ncu -k regex:conv1x1_tc_kernel`.
    python tools/profile_conv.py CIN COUT HW RES(0|1) [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quantized_neural_nets_b200._lib import lib, launch

cin, cout, hw, with_res = (int(v) for v in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 256
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.relu(torch.randn(B, cin, hw, hw, device=dev, generator=g))
w = torch.randn(cout, cin, device=dev, generator=g) * 0.05
alpha = torch.rand(cout, device=dev, generator=g) + 0.5
beta = torch.randn(cout, device=dev, generator=g) * 0.1
res = torch.randn(B, cout, hw, hw, device=dev, generator=g) if with_res else None
out = torch.empty(B, cout, hw, hw, device=dev)
ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(cout, cin), dtype=torch.uint8, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    if it == 3:
        a.record()
    launch(lib.gpfq_conv1x1_bn_act_f32, x, hw * hw, w, res, alpha, beta, out, B, cin, cout, hw * hw, 0.0, float("inf"), ws,
           ws.numel())
b.record()
torch.cuda.synchronize()
print(f"{cin}->{cout} @{hw} res={with_res} B={B}: {a.elapsed_time(b) / 3:.3f} ms per launch")
