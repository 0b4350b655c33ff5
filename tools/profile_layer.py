"""Runs the solver on one synthetic layer shape (for ncu / quick timing).
usage: python tools/profile_layer.py N d m [reps] [solver]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quantized_neural_nets_b200 as qb
from quantized_neural_nets_b200 import _lib
from quantized_neural_nets_b200.step_algorithm import quantize_layer_impl

N, d, m = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
solver = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
W = torch.randn(N, d, device=dev, generator=g) * 0.05
ld = (m + 3) // 4 * 4
Xfm = torch.relu(torch.randn(d, ld, device=dev, generator=g))
Xqfm = torch.relu(Xfm + 0.02 * torch.randn(d, ld, device=dev, generator=g))
X, Xq = Xfm[:, :m].t(), Xqfm[:, :m].t()
for r in range(reps):
    torch.cuda.synchronize()
    if not os.environ.get('NO_PROFILE'): _lib.profile_begin()
    t0 = time.perf_counter()
    Q, e2, r2 = quantize_layer_impl(W, X, Xq, m, 1.16 / 8, 8, 1, None, 0.1, 1, False, dev, return_partials=True,
                                    solver=solver)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    p = _lib.profile_end() if not os.environ.get('NO_PROFILE') else dict(sweep_ms=0.0, sweep_launches=0, sweep_fp32_instr=0.0)
    units = float(N) * d * m
    print(f"rep {r}: {dt*1e3:.3f} ms total, sweep {p['sweep_ms']:.3f} ms in {p['sweep_launches']} launches, "
          f"{units/dt:.3e} w*s/s, sweep fp32 {p['sweep_fp32_instr']/max(p['sweep_ms'],1e-9)/1e6:.1f} Ginstr/s, "
          f"rel {float((e2.sum()/r2.sum()).sqrt()):.5f}")
