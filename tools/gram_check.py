"""Accuracy / speed of the Gram-matrix kernels (fp64 SIMT and tcgen05 split-TF32) against torch float64.
usage: python tools/gram_check.py d m [solver ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_neural_nets_b200 import _lib
from quantized_neural_nets_b200._lib import lib, launch

d, m = int(sys.argv[1]), int(sys.argv[2])
solvers = [int(s) for s in sys.argv[3:]] or [2, 1]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ld = (m + 3) // 4 * 4
X = torch.relu(torch.randn(d, ld, device=dev, generator=g)); X[:, m:] = 0
Xq = torch.relu(X + 0.02 * torch.randn(d, ld, device=dev, generator=g)); Xq[:, m:] = 0
Xd, Xqd = X[:, :m].double(), Xq[:, :m].double()
ref = {"GT": Xd @ Xqd.T, "H": Xqd @ Xqd.T, "A": Xd @ Xd.T}
ldg = (d + 63) // 64 * 64
for sv in solvers:
    nbytes = lib.gpfq_workspace_bytes(sv, 1, d, m)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = [torch.full((ldg, ldg), float("nan"), dtype=torch.float64, device=dev) for _ in range(3)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        launch(lib.gpfq_gram_f32, sv, X, Xq, ld, d, m, out[0], out[1], out[2], ws, nbytes)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    flops = 2.0 * d * d * m * 3
    line = f"solver {sv} d={d} m={m}: {dt*1e3:.3f} ms ({flops/dt/1e12:.1f} algorithmic TFLOP/s, workspace {nbytes/1e6:.0f} MB)"
    for name, o in zip(("GT", "H", "A"), out):
        err = (o[:d, :d] - ref[name]).abs()
        line += f" | {name}: max rel err {float((err / ref[name].abs().clamp_min(1e-30)).max()):.2e} rel fro {float(err.norm() / ref[name].norm()):.2e}"
    print(line, flush=True)
