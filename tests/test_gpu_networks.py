"""Teacher-forced per-layer parity on real networks (random-init torchvision models, Gaussian
images): every layer's (W, X, X~) as produced by the CUDA orchestrator is also handed to the CPU
oracle, and the two solutions are compared level by level.  Needs a B200: -m gpu."""
import numpy as np
import pytest
import torch
import torchvision

from oracle import gpfq_oracle as orc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def run_teacher_forced(name, batch, bits=4, reg=None, lam=0.1, max_layers=None):
    import quantized_neural_nets_b200 as qb
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = getattr(torchvision.models, name)(weights=None).eval().to(DEV)
    layers = []
    qb.extract_layers(model, layers)
    g = torch.Generator().manual_seed(1)
    loader = [(torch.randn(batch, 3, 224, 224, generator=g), None) for _ in layers]
    np.random.seed(0)
    qnn = qb.QuantizeNeuralNet(model, name, batch, loader, bits, bits, [], 1.16, 1.16, 1, 1, reg, lam, 0.25, False, DEV)
    K = 2 ** (bits - 1)
    report = []
    for i, layer in enumerate(qnn.analog_network_layers[:max_layers]):
        X, Xq = qnn._populate_linear_layer_input(i)
        W = layer.weight.data.view(layer.weight.shape[0], -1)
        groups = getattr(layer, "groups", 1)
        Q, err, rel, _, _ = qb.StepAlgorithm._quantize_layer(W, X, Xq, X.shape[0], 1.16 / K, K, 1, reg, lam, groups,
                                                            False, DEV)
        Wc, Xc, Xqc = W.cpu(), X.cpu().contiguous(), Xq.cpu().contiguous()
        Qo, erro, relo, _, _ = orc.quantize_layer(Wc, Xc, Xqc, Xc.shape[0], 1.16 / K, K, 1, reg, lam, groups, False)
        delta = orc.layer_step_size(Wc, 1.16 / K, K, 1, reg, lam)
        lv, lvo = orc.level_index(Q.cpu(), delta, reg, lam), orc.level_index(Qo, delta, reg, lam)
        report.append((i, tuple(W.shape), X.shape[0], (lv == lvo).float().mean().item(), rel.item(), float(relo)))
        qnn.quantized_network_layers[i].weight.data = Q.reshape(layer.weight.shape).float()
    return report


def check(report):
    for i, shape, m, agree, rel, relo in report:
        assert agree >= 0.999, f"layer {i} {shape} m={m}: level agreement {agree}"
        assert abs(rel - relo) <= 1e-3 * relo, f"layer {i}: rel err {rel} vs oracle {relo}"
    total = sum(a * s[0] * s[1] for _, s, _, a, _, _ in report) / sum(s[0] * s[1] for _, s, _, a, _, _ in report)
    return total


def test_resnet18_every_layer_teacher_forced():
    report = run_teacher_forced("resnet18", batch=8)
    assert len(report) == 21
    assert check(report) >= 0.9995


def test_alexnet_soft_threshold_teacher_forced():
    # AlexNet exercises d = 9216 (288 feature blocks) and the L1 (soft-threshold) alphabet
    report = run_teacher_forced("alexnet", batch=4, reg="L1", lam=1e-4)
    assert len(report) == 8
    assert check(report) >= 0.9995
