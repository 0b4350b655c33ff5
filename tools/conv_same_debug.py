import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_neural_nets_b200._lib import lib, launch
dev = torch.device("cuda:0")
B, C, N, H, W = 2, 64, 64, 56, 56
x = torch.randn(B, C, H, W, device=dev); w = torch.randn(N, C, 3, 3, device=dev) * 0.05
out = torch.zeros(B, N, H, W, device=dev)
ws = torch.empty(lib.gpfq_conv_same_workspace_bytes(N, C, 3, 3, B, H, W), dtype=torch.uint8, device=dev)
launch(lib.gpfq_conv_same_bn_act_f32, x, w, None, None, None, out, B, C, N, H, W, 3, 3, -float("inf"), float("inf"), ws, ws.numel())
torch.cuda.synchronize()
ref = torch.nn.functional.conv2d(x, w, padding=1)
print("debug", os.environ.get("GPFQ_CONV_DEBUG"), "ok; rel diff", float((out - ref).norm() / ref.norm()))
