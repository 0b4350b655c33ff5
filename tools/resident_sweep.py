"""Times the direct solver's launch structures on a list of layer shapes (tuning aid for make_plan's cost model).
usage: python tools/resident_sweep.py [N,d,m ...]      (default: the launch-bound ResNet-50 / AlexNet shapes)
For every shape: multi-launch (GPFQ_RESIDENT=0), the library's own choice, and the resident kernel forced with
each (cluster, TN) combination.  Times are CUDA-event best-of-3 of the whole solve call, in ms."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_neural_nets_b200 import _lib
from quantized_neural_nets_b200.step_algorithm import quantize_layer_impl

DEFAULT = ["256,2304,1792", "512,4608,1792", "512,4608,768", "512,2048,3328", "1000,2048,256", "128,1152,6656",
           "256,2304,6656", "256,1024,12800", "512,1024,12800", "2048,512,3328", "64,576,23296", "128,1152,23296",
           "4096,9216,256", "4096,4096,256"]
shapes = [tuple(int(v) for v in s.split(",")) for s in (sys.argv[1:] or DEFAULT)]
dev = torch.device("cuda:0")
KEYS = ("GPFQ_RESIDENT", "GPFQ_RESIDENT_CLUSTER", "GPFQ_RESIDENT_TN")


DELTA = torch.tensor(0.02, device=dev)


def run(W, X, Xq, m, env):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    best, rel, res = 1e9, 0.0, False
    for r in range(4):
        _lib.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Q, e2, r2 = quantize_layer_impl(W, X, Xq, m, 1.16 / 8, 8, 1, None, 0.1, 1, False, dev, return_partials=True,
                                        solver=0, delta=DELTA)
        e1.record()
        torch.cuda.synchronize()
        p = _lib.profile_end()
        res = p["resident_launches"] > 0
        if r:
            best = min(best, e0.elapsed_time(e1))
        rel = float((e2.sum() / r2.sum()).sqrt())
    return best, rel, res, Q


for (N, d, m) in shapes:
    g = torch.Generator(device=dev).manual_seed(0)
    W = torch.randn(N, d, device=dev, generator=g) * 0.05
    ld = (m + 3) // 4 * 4
    Xfm = torch.relu(torch.randn(d, ld, device=dev, generator=g))
    Xqfm = torch.relu(Xfm + 0.02 * torch.randn(d, ld, device=dev, generator=g))
    X, Xq = Xfm[:, :m].t(), Xqfm[:, :m].t()
    t_multi, rel0, _, Q0 = run(W, X, Xq, m, {"GPFQ_RESIDENT": "0"})
    t_auto, _, res_auto, _ = run(W, X, Xq, m, {})
    line = [f"{N}x{d}x{m}: multi {t_multi:.3f}  auto {t_auto:.3f}{'(res)' if res_auto else ''} |"]
    for cs in (1, 2, 4, 8, 16):
        for tn in (32, 16):
            t, rel, res, Q = run(W, X, Xq, m, {"GPFQ_RESIDENT": "1", "GPFQ_RESIDENT_CLUSTER": str(cs),
                                               "GPFQ_RESIDENT_TN": str(tn)})
            if not res:
                continue
            same = float((Q == Q0).float().mean())
            line.append(f"c{cs}t{tn} {t:.3f}" + ("" if same == 1.0 else f"[agree {same:.5f}]"))
    print("  ".join(line), flush=True)
