"""Diagnose level mismatches between the CUDA path and the CPU oracle on a real network layer by layer:
for every neuron whose path diverges, report the float64 decision margin at the first divergence, and compare both
paths against a float64 run of the same algorithm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torchvision
import quantized_neural_nets_b200 as qb
from oracle import gpfq_oracle as orc

name = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
DEV = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
model = getattr(torchvision.models, name)(weights=None).eval().to(DEV)
layers = []; qb.extract_layers(model, layers)
g = torch.Generator().manual_seed(1)
loader = [(torch.randn(batch, 3, 224, 224, generator=g), None) for _ in layers]
np.random.seed(0)
qnn = qb.QuantizeNeuralNet(model, name, batch, loader, 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25, False, DEV)
K = 8

def f64_path(W, X, Xq, delta):
    W, X, Xq = W.double(), X.double(), Xq.double(); d = float(delta)
    N, dd = W.shape; u = torch.zeros(N, X.shape[0], dtype=torch.float64); Q = torch.zeros(N, dd, dtype=torch.float64)
    for t in range(dd):
        u += torch.outer(W[:, t], X[:, t]); n = (Xq[:, t] ** 2).sum()
        a = u @ Xq[:, t] / n if n > 0 else torch.zeros(N, dtype=torch.float64)
        q = torch.sign(a) * d * torch.clamp(torch.floor(a / d + 0.5).abs(), max=K)
        Q[:, t] = q; u -= torch.outer(q, Xq[:, t])
    return Q

for i, layer in enumerate(qnn.analog_network_layers):
    X, Xq = qnn._populate_linear_layer_input(i)
    W = layer.weight.data.view(layer.weight.shape[0], -1)
    Wc, Xc, Xqc = W.cpu(), X.cpu().contiguous(), Xq.cpu().contiguous()
    if os.environ.get("CPU_DELTA"):
        dl = orc.layer_step_size(Wc, 1.16 / K, K, 1, None, 0.1)
        Q = torch.zeros_like(W); U = torch.zeros(W.shape[0], X.shape[0], device=DEV)
        qb.StepAlgorithm._quantization(W, Q, U, X, Xq, qb.StepAlgorithm._msq, dl.to(DEV), K, 0.1)
    elif os.environ.get("SOLVER"):
        from quantized_neural_nets_b200 import step_algorithm as sa
        sv = int(os.environ["SOLVER"])
        if not sa.gram_eligible(W.shape[0], W.shape[1], X.shape[0]):
            sv = 0
        Q, e2, r2 = sa.quantize_layer_impl(W, X, Xq, X.shape[0], 1.16 / K, K, 1, None, 0.1, 1, False, DEV, solver=sv,
                                           return_partials=True)
        print(f"   [solver {sv}]", end="")
    else:
        Q, err, rel, _, _ = qb.StepAlgorithm._quantize_layer(W, X, Xq, X.shape[0], 1.16 / K, K, 1, None, 0.1, 1, False, DEV)
    Qo, erro, relo, _, _ = orc.quantize_layer(Wc, Xc, Xqc, Xc.shape[0], 1.16 / K, K, 1, None, 0.1, 1, False)
    delta = orc.layer_step_size(Wc, 1.16 / K, K, 1, None, 0.1)
    lv, lvo = orc.level_index(Q.cpu(), delta), orc.level_index(Qo, delta)
    diff = lv != lvo
    line = f"layer {i:2d} {tuple(W.shape)} m={X.shape[0]} agree={1 - diff.float().mean().item():.6f}"
    if diff.any():
        Q64 = f64_path(Wc, Xc, Xqc, delta); lv64 = orc.level_index(Q64, delta)
        margin = orc.exact_decision_margin(Wc, Xc, Xqc, Qo, delta, K)
        firsts = [(n, int(diff[n].nonzero()[0])) for n in diff.any(1).nonzero().flatten().tolist()]
        line += f" | neurons diverged {len(firsts)} first-divergence margins {[f'{float(margin[n, t]):.1e}' for n, t in firsts[:8]]}"
        line += f" | vs f64 path: cuda {1 - (lv != lv64).float().mean().item():.6f} oracle {1 - (lvo != lv64).float().mean().item():.6f}"
    print(line, flush=True)
    qnn.quantized_network_layers[i].weight.data = Q.reshape(layer.weight.shape).float()
