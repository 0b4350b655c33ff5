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


def run_teacher_forced(name, batch, bits=4, reg=None, lam=0.1, max_layers=None, solver=None, grouped=False):
    import quantized_neural_nets_b200 as qb
    from quantized_neural_nets_b200 import step_algorithm as sa
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
    n_batched = []
    for i, layer in enumerate(qnn.analog_network_layers[:max_layers]):
        X, Xq = qnn._populate_linear_layer_input(i)
        W = layer.weight.data.view(layer.weight.shape[0], -1)
        groups = getattr(layer, "groups", 1)
        if grouped and groups > 1 and sa.grouped_eligible(groups, W.shape[1], X.shape[0]):
            Q, err, rel, _, _ = sa.quantize_layer_impl(W, X, Xq, X.shape[0], 1.16 / K, K, 1, reg, lam, groups, False, DEV,
                                                       solver=sa.GROUPED)       # all groups in one batched solve
            n_batched.append(i)
        elif solver is None:
            Q, err, rel, _, _ = qb.StepAlgorithm._quantize_layer(W, X, Xq, X.shape[0], 1.16 / K, K, 1, reg, lam, groups,
                                                                False, DEV)
        else:   # Gram solver where its shape rule applies, direct elsewhere
            sv = solver if (groups == 1 and sa.gram_eligible(W.shape[0], W.shape[1], X.shape[0])) else 0
            Q, e2, r2 = sa.quantize_layer_impl(W, X, Xq, X.shape[0], 1.16 / K, K, 1, reg, lam, groups, False, DEV,
                                               solver=sv, return_partials=True)
            err, rel, _, _ = sa.reduce_errors(e2, r2, groups)
        Wc, Xc, Xqc = W.cpu(), X.cpu().contiguous(), Xq.cpu().contiguous()
        Qo, erro, relo, _, _ = orc.quantize_layer(Wc, Xc, Xqc, Xc.shape[0], 1.16 / K, K, 1, reg, lam, groups, False)
        delta = orc.layer_step_size(Wc, 1.16 / K, K, 1, reg, lam)
        lv, lvo = orc.level_index(Q.cpu(), delta, reg, lam), orc.level_index(Qo, delta, reg, lam)
        diff = lv != lvo
        worst_margin = 0.0
        if diff.any() and groups == 1:   # every neuron's FIRST divergence must sit at a rounding tie
            margin = orc.exact_decision_margin(Wc, Xc, Xqc, Qo, delta, K, reg, lam)
            for n in diff.any(dim=1).nonzero().flatten().tolist():
                worst_margin = max(worst_margin, float(margin[n, int(diff[n].nonzero()[0])]))
        report.append((i, tuple(W.shape), X.shape[0], 1.0 - diff.float().mean().item(), rel.item(), float(relo),
                       worst_margin))
        qnn.quantized_network_layers[i].weight.data = Q.reshape(layer.weight.shape).float()
    run_teacher_forced.batched_layers = list(n_batched)
    return report


def check(report):
    """BASELINE.json's gate: >= 99.9 % identical levels (weighted over the network's weights), differences only
    at rounding ties, per-layer relative error within 1e-3 relative.  A single tie flip diverts one whole neuron
    (1/N of a layer), so individual small-N layers are only required to stay above 99 %."""
    for i, shape, m, agree, rel, relo, margin in report:
        assert agree >= 0.99, f"layer {i} {shape} m={m}: level agreement {agree}"
        if m >= shape[1]:   # the float64 margin is only meaningful where the layer is well posed (m >= d)
            assert margin < 2e-4, f"layer {i} {shape} m={m}: first divergence {margin} away from a rounding tie"
        if np.isnan(relo):      # a group whose inputs are all zero (dead ReLU channel): 0 / 0 in the reference too
            assert np.isnan(rel), f"layer {i}: rel err {rel} vs oracle nan"
        else:
            assert abs(rel - relo) <= 1e-3 * relo, f"layer {i}: rel err {rel} vs oracle {relo}"
    weights = sum(s[0] * s[1] for _, s, *_ in report)
    return sum(r[3] * r[1][0] * r[1][1] for r in report) / weights


def test_resnet18_every_layer_teacher_forced():
    report = run_teacher_forced("resnet18", batch=8)
    assert len(report) == 21
    assert check(report) >= 0.999


def test_resnet50_gram_tcgen05_teacher_forced():
    """ResNet-50 with the tcgen05 Gram solver on every layer its shape rule selects (the 1x1 expand and
    downsample convolutions), direct solver elsewhere."""
    report = run_teacher_forced("resnet50", batch=16, solver=1)
    assert len(report) == 54
    assert check(report) >= 0.999


def test_alexnet_soft_threshold_teacher_forced():
    # AlexNet exercises d = 9216 (288 feature blocks) and the L1 (soft-threshold) alphabet
    report = run_teacher_forced("alexnet", batch=4, reg="L1", lam=1e-4)
    assert len(report) == 8
    assert check(report) >= 0.999


def test_mobilenet_v2_depthwise_layers_batched_teacher_forced():
    """MobileNetV2 (17 depthwise 3x3 convolutions, up to 960 groups of one neuron and 9 features): the batched
    grouped solver against the oracle's loop over groups (step_algorithm.py:221-247), every layer teacher-forced."""
    report = run_teacher_forced("mobilenet_v2", 4, grouped=True)
    assert len(report) == 53 and len(run_teacher_forced.batched_layers) == 17
    assert check(report) >= 0.999


def _quantize(model, loader, calibration, batch, ignore=(), reg=None, lam=0.1):
    import quantized_neural_nets_b200 as qb
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    np.random.seed(3)
    qnn = qb.QuantizeNeuralNet(model, "net", batch, loader, 4, 4, list(ignore), 1.16, 1.16, 1, 1, reg, lam, 0.25, False,
                               DEV, calibration=calibration, solver=0)
    qnn.quantize_network()
    return qnn


@pytest.mark.parametrize("name", ["tiny", "resnet18"])
def test_reuse_calibration_equals_reference_schedule_on_a_repeated_batch(name):
    """calibration='reuse' (one analog pass + one quantizing pass) must compute exactly what
    the reference's per-layer schedule (quantize_neural_net.py:117-214) computes when the loader yields the same
    batch for every layer: same conv patch draws, same X / X~, hence bit-identical Q and errors."""
    import copy
    import golden_cases as gc
    if name == "tiny":
        model, batch, size, ignore = gc.tiny_cnn(0).to(DEV), 6, 16, [0]
    else:
        torch.manual_seed(0)
        model, batch, size, ignore = torchvision.models.resnet18(weights=None).eval().to(DEV), 8, 224, []
    images = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(7))
    n_layers = len(_quantize(copy.deepcopy(model), [(images, None)], "reuse", batch, ignore).quantized_network_layers)
    fresh = _quantize(copy.deepcopy(model), [(images, None)] * n_layers, "fresh", batch, ignore)
    reuse = _quantize(copy.deepcopy(model), [(images, None)], "reuse", batch, ignore)
    assert len(fresh.layer_log) == len(reuse.layer_log) == n_layers - len(ignore)
    for lf, lr in zip(fresh.quantized_network_layers, reuse.quantized_network_layers):
        assert torch.equal(lf.weight.data, lr.weight.data)
    for (i, e, r), (j, e2, r2) in zip(fresh.layer_log, reuse.layer_log):
        assert i == j and float(e) == float(e2) and float(r) == float(r2)
    with torch.no_grad():
        assert torch.equal(fresh.quantized_network(images.to(DEV)), reuse.quantized_network(images.to(DEV)))


@pytest.mark.parametrize("shape,lo,hi,with_res", [((3, 5, 7, 7), 0.0, float("inf"), False), ((2, 8, 12, 12), 0.0, 6.0, False),
                                                  ((2, 8, 12, 12), 0.0, float("inf"), True),
                                                  ((4, 3, 6, 10), -float("inf"), float("inf"), False),
                                                  ((1, 16, 9, 9), 0.0, float("inf"), True)])
def test_fused_bn_act_kernel_is_exact(shape, lo, hi, with_res):
    """gpfq_bn_act_f32 against the same expression in separately rounded torch ops (x * alpha + beta (+ r), clamp)."""
    from quantized_neural_nets_b200.forward_fusion import FusedBNAct
    g = torch.Generator().manual_seed(3)
    C = shape[1]
    bn = torch.nn.BatchNorm2d(C).eval()
    bn.running_mean = torch.randn(C, generator=g)
    bn.running_var = torch.rand(C, generator=g) + 0.3
    bn.weight.data = torch.randn(C, generator=g)
    bn.bias.data = torch.randn(C, generator=g)
    bn = bn.to(DEV)
    x = torch.randn(shape, generator=g).to(DEV)
    x[0, 0, 0, 0] = float("nan")
    res = torch.randn(shape, generator=g).to(DEV) if with_res else None
    got = FusedBNAct(bn, lo, hi)(x, res)
    alpha = (1.0 / torch.sqrt(bn.running_var + bn.eps)) * bn.weight.data
    beta = bn.bias.data - bn.running_mean * alpha
    want = x * alpha[None, :, None, None] + beta[None, :, None, None]
    if with_res:
        want = want + res
    want = torch.clamp(want, min=lo, max=hi)
    assert torch.equal(torch.nan_to_num(got, nan=123.0), torch.nan_to_num(want, nan=123.0))
    assert torch.isnan(got[0, 0, 0, 0])
    # and it is PyTorch's own batch norm up to rounding
    ref = bn(x) if res is None else bn(x) + res
    ref = torch.clamp(ref, min=lo, max=hi)
    assert torch.allclose(torch.nan_to_num(got), torch.nan_to_num(ref), rtol=1e-5, atol=1e-6)


def test_fused_forward_matches_plain_forward_and_keeps_hooks():
    from quantized_neural_nets_b200.forward_fusion import fuse_inference_forward
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = torchvision.models.resnet18(weights=None).eval().to(DEV)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    fused, sites = fuse_inference_forward(model)
    assert sites == 20
    x = torch.randn(8, 3, 224, 224, device=DEV)
    seen = []
    handle = model.layer3[0].conv1.register_forward_hook(lambda m, i, o: seen.append(i[0].shape))
    with torch.no_grad():
        a, b = model(x), fused(x)
    handle.remove()
    assert len(seen) == 2 and (a - b).norm() <= 1e-5 * a.norm()
    model.fc.weight.data.zero_()                       # the fused callable shares the modules' parameters
    with torch.no_grad():
        assert torch.equal(fused(x), model.fc.bias.data.expand(8, -1))


def test_quantize_network_with_fused_forward():
    """fuse_forward=True changes how the calibration activations are computed (one elementwise pass instead of
    cuDNN batch norm + add + ReLU), not what is computed: the quantized network agrees with the unfused run up to
    rare tie flips."""
    import copy
    import golden_cases as gc
    torch.backends.cudnn.allow_tf32 = False
    results = []
    for fuse in (False, True):
        model = copy.deepcopy(gc.bn_cnn(0)).to(DEV)
        np.random.seed(5)
        import quantized_neural_nets_b200 as qb
        qnn = qb.QuantizeNeuralNet(model, "bn", 6, gc.image_batches(4, 6, 8, 51), 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1,
                                   0.5, False, DEV, fuse_forward=fuse)
        assert bool(qnn._fused) == fuse
        results.append(qnn.quantize_network())
    for la, lb in zip(results[0].modules(), results[1].modules()):
        if isinstance(la, (torch.nn.Conv2d, torch.nn.Linear)):
            assert (la.weight.data == lb.weight.data).float().mean().item() >= 0.99
    probe = gc.image_batches(1, 5, 8, 52)[0][0].to(DEV)
    with torch.no_grad():
        a, b = results[0](probe), results[1](probe)
    assert (a - b).norm() <= 2e-2 * a.norm()


def test_pointwise_convs_as_gemm_match_cudnn_and_keep_hooks():
    from quantized_neural_nets_b200.forward_fusion import pointwise_convs_as_gemm
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).eval().to(DEV)
    x = torch.randn(8, 3, 224, 224, device=DEV)
    conv = torch.nn.Conv2d(96, 40, 1, bias=True).to(DEV)
    y = torch.randn(5, 96, 9, 13, device=DEV)
    with torch.no_grad():
        want, want_b = model(x), conv(y)
    seen = []
    handle = model.layer2[1].conv3.register_forward_hook(lambda m, i, o: seen.append(tuple(o.shape)))
    with pointwise_convs_as_gemm(model, conv) as n, torch.no_grad():
        assert n == 34
        got, got_b = model(x), conv(y)
    handle.remove()
    assert seen == [(8, 512, 28, 28)]
    # the 1x1 layers now run on the tensor cores in split-TF32 (fp32-level, but not cuDNN's rounding): 50 layers deep
    assert (got - want).norm() <= 1e-5 * want.norm() and torch.allclose(got_b, want_b, rtol=1e-4, atol=1e-5)
