"""Role timeline of the tensor-core convolution: where does a stage's cycle go?  Needs a library built with
-DGPFQ_CONV_TRACE (python tools/conv_trace.py --build; rebuild normally afterwards).  CTA 0 stamps clock64() at the
hand-over points of its first 256 k-blocks; printed are the median intervals (cycles) in steady state:
    python tools/conv_trace.py CIN COUT HW RES(0|1)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if "--build" in sys.argv:
    import importlib.util
    spec = importlib.util.spec_from_file_location("gpfq_build", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "quantized_neural_nets_b200", "build.py"))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)
    build.NVCC_FLAGS.append("-DGPFQ_CONV_TRACE")
    build.build(force=True)
    sys.exit(0)

import numpy as np
import torch
from quantized_neural_nets_b200._lib import lib, launch

dev = torch.device("cuda:0")
B = 256
g = torch.Generator(device=dev).manual_seed(0)
L = 256
for spec in sys.argv[1:]:
    cin, cout, hw, with_res = (int(v) for v in spec.split(","))
    x = torch.relu(torch.randn(B, cin, hw, hw, device=dev, generator=g))
    w = torch.randn(cout, cin, device=dev, generator=g) * 0.05
    alpha = torch.rand(cout, device=dev, generator=g) + 0.5
    beta = torch.randn(cout, device=dev, generator=g) * 0.1
    res = torch.randn(B, cout, hw, hw, device=dev, generator=g) if with_res else None
    out = torch.empty(B, cout, hw, hw, device=dev)
    ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(cout, cin), dtype=torch.uint8, device=dev)
    trace = torch.zeros(8 * L, dtype=torch.int64, device=dev)
    os.environ["GPFQ_CONV_TRACE_PTR"] = str(trace.data_ptr())
    for _ in range(3):
        launch(lib.gpfq_conv1x1_bn_act_f32, x, hw * hw, w, res, alpha, beta, out, B, cin, cout, hw * hw, 0.0, float("inf"), ws,
               ws.numel())
    torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(8, L).astype(np.int64)
    nkb = (cin + 31) // 32
    lo, hi = 3 * nkb + 8, L - 8          # steady state
    sl = slice(lo, hi)
    med = lambda a: int(np.median(a))
    print(f"{cin} -> {cout} @ {hw}{' + res' if with_res else ''}: k-blocks per tile {nkb}")
    print(f"  k-block period (stage free -> next stage free)       {med(np.diff(t[0][sl]))}")
    print(f"  TMA: stage free -> tile landed                       {med(t[1][sl] - t[0][sl])}")
    print(f"  split: landed -> planes ready                        {med(t[2][sl] - t[1][sl])}")
    print(f"  MMA warp: planes ready -> ready to issue             {med(t[3][sl] - t[2][sl])}")
    print(f"  MMA issue (12 MMAs + 2 commits)                      {med(t[4][sl] - t[3][sl])}")
    print(f"  tensor core: issued -> accumulator complete (drain)  {med(t[5][sl] - t[4][sl])}")
    print(f"  drain: accumulator complete -> drained               {med(t[6][sl] - t[5][sl])}")
    print(f"  stage recycle: MMAs issued(it) -> stage free(it+3)   {med(t[0][lo + 3:hi] - t[4][lo:hi - 3])}")
    ep = t[7][nkb - 1::nkb]
    last_drain = t[6][nkb - 1::nkb]
    n = min(len(ep), len(last_drain)) - 1
    print(f"  epilogue: last drain -> tile stored                  {med(ep[3:n] - last_drain[3:n])}")
    print(f"  tile period (drain warp)                             {med(np.diff(ep[3:n]))}")
    first = t[5][0::nkb]                 # accumulator-complete stamp of each tile's first k-block
    print(f"  tile stored -> next tile's first accumulator taken   {med(first[4:n + 1] - ep[3:n])}")
    if nkb > 1:
        inner = (t[5][1:] - t[6][:-1]).reshape(-1)
        idx = np.array([i for i in range(lo, hi) if (i + 1) % nkb != 0])
        print(f"  drained(kb) -> accumulator(kb+1) taken, same tile    {med(inner[idx])}")
    print(f"  MMA warp: issue(it) done -> ready for it+1           {med(t[3][lo + 1:hi] - t[4][lo:hi - 1])}")
